"""One DeepLab training step (bs=16, 513x513) between cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import NUM_CLASSES, synth_batch  # noqa: E402
from zs3_b200.modeling.deeplab import DeepLab  # noqa: E402
from zs3_b200.parallel import DataParallelTrainer  # noqa: E402
from zs3_b200.utils.loss import SegmentationLosses  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(1)
dev = torch.device("cuda", 0)
model = DeepLab(num_classes=NUM_CLASSES, output_stride=16, sync_bn=True, pretrained=False).to(dev).train()
trainer = DataParallelTrainer(model, SegmentationLosses(cuda=True).build_loss("ce"))
img, lab = synth_batch(batch, 513, 7, device=dev)
for _ in range(2):
    trainer.train_step(img, lab)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    trainer.train_step(img, lab)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", steps, "step(s)")
