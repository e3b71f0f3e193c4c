"""Why is N-GPU data-parallel efficiency below 1?  Every rank of one box first trains ALONE (no collective, its own
CUDA graph: the pure per-GPU step time while all N GPUs are busy), then the same ranks train data-parallel.  Prints the
per-rank independent step times, their max, and the data-parallel step time.
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/scale_probe.py [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import NUM_CLASSES, synth_batch  # noqa: E402
from zs3_b200.modeling.deeplab import DeepLab  # noqa: E402
from zs3_b200.parallel import DataParallelTrainer, init_distributed  # noqa: E402
from zs3_b200.utils.loss import SegmentationLosses  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rank, local, world = init_distributed()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
torch.manual_seed(1)


SEG = {}


def run(world_size):
    model = DeepLab(num_classes=NUM_CLASSES, output_stride=16, sync_bn=True, pretrained=False).to(dev).train()
    tr = DataParallelTrainer(model, SegmentationLosses(cuda=True).build_loss("ce"), world_size=world_size, use_cuda_graph=True)
    img, lab = synth_batch(16, 513, 7 + rank, device=dev)
    for _ in range(4):
        tr.train_step(img, lab)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tr.segment_events = []
    e0.record()
    for _ in range(steps):
        tr.train_step(img, lab)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    seg = [[a.elapsed_time(b) for a, b in zip(ev, ev[1:])] for ev in tr.segment_events]
    SEG[world_size] = [round(sum(s[k] for s in seg) / len(seg), 3) for k in range(3)]
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    del tr, model
    torch.cuda.empty_cache()
    return [float(o.item()) for o in out]


alone = run(1)
together = run(world)
if rank == 0:
    print(json.dumps({"n_gpus": world, "steps": steps,
                      "independent_replicas_ms_per_step_by_rank": [round(v, 3) for v in alone],
                      "independent_max": round(max(alone), 3), "independent_min": round(min(alone), 3),
                      "data_parallel_ms_per_step_by_rank": [round(v, 3) for v in together],
                      "data_parallel_max": round(max(together), 3),
                      "dp_over_slowest_replica": round(max(together) / max(alone), 4),
                      "dp_over_fastest_replica": round(max(together) / min(alone), 4),
                      "rank0_segments_ms [graph 1 | graph 2 (tail) | after the tail]": {"alone": SEG.get(1), "data_parallel": SEG.get(world)}}))
dist.barrier()
dist.destroy_process_group()
