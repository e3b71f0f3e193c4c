#!/usr/bin/env python
"""BASELINE.json configs[2]: one ZS3Net step-2 iteration (train_pascal_GMMN.py:152-268) -- DeepLab feature extraction
under no_grad, per-(image, class) generator updates, classifier (pred_conv) update -- bs=16, 513x513, 21 classes.

Times `ZS3Step` (module-by-module, as the unchanged reference trainer drives the modules) and `ZS3StepFused` (label
work list on the device + the fused generator-update kernel) with CUDA events, and the fused kernel alone on the
step's work list.  Writes one JSON object (stdout and --out).

    python tools/step2_bench.py --steps 5 --warmup 3 --out gpurun_out/step2_bench.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

UNSEEN = [10, 14]          # Pascal-VOC 2-unseen split of the ZS3 paper (cow, motorbike)
C = 21


def synth_labels(n, hw, seed, unseen_fraction=0.25):
    """blocky label maps with 2-5 classes per image; ~25 % of the images hold an unseen class; 2 % ignore pixels"""
    g = torch.Generator().manual_seed(seed)
    seen = [c for c in range(C) if c not in UNSEEN]
    lab = torch.zeros(n, hw, hw)
    for i in range(n):
        k = int(torch.randint(2, 6, (1,), generator=g))
        cls = [seen[j] for j in torch.randperm(len(seen), generator=g)[:k].tolist()]
        if torch.rand(1, generator=g).item() < unseen_fraction:
            cls[-1] = UNSEEN[int(torch.randint(0, len(UNSEEN), (1,), generator=g))]
        grid = torch.randint(0, k, (8, 8), generator=g)
        cell = (hw + 7) // 8
        lab[i] = torch.tensor(cls, dtype=torch.float32)[grid].repeat_interleave(cell, 0).repeat_interleave(cell, 1)[:hw, :hw]
    lab[torch.rand(n, hw, hw, generator=g) < 0.02] = 255
    return lab


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=513)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="")
    ap.add_argument("--skip-unfused", action="store_true")
    args = ap.parse_args()

    from zs3.modeling.deeplab import DeepLab
    from zs3.modeling.gmmn import GMMNnetwork
    from zs3.utils.loss import GMMNLoss, SegmentationLosses
    from zs3_b200 import _lib as L
    from zs3_b200.step2 import ZS3Step, ZS3StepFused

    torch.manual_seed(1)
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    seen = [c for c in range(C) if c not in UNSEEN]
    image = torch.randn(args.batch, 3, args.size, args.size, device=dev)
    target = synth_labels(args.batch, args.size, seed=2).to(dev)
    emb_table = (torch.randn(C, 300, generator=torch.Generator().manual_seed(8)) * 0.06).to(dev)
    # per-pixel embedding map as the reference's dataloader builds it (dataloaders/datasets/base.py:45-51): 5 GB
    embedding = emb_table[target.clamp(max=C - 1).long()].permute(0, 3, 1, 2).contiguous()

    def build(step_cls, **kw):
        torch.manual_seed(1)
        model = DeepLab(num_classes=C, sync_bn=True, freeze_bn=False, pretrained=False)
        model = torch.nn.DataParallel(model.cuda(), device_ids=[0])
        model.train()
        gen = GMMNnetwork(300, 300, 256, 256).cuda().train()
        cw = torch.ones(C)
        cw[UNSEEN] = 100.0
        crit = SegmentationLosses(weight=cw.cuda(), cuda=True).build_loss("ce")
        crit_g = GMMNLoss(sigma=[2, 5, 10, 20, 40, 80], cuda=True).build_loss()
        opt = torch.optim.SGD([{"params": model.module.get_1x_lr_params(), "lr": 0.007},
                               {"params": model.module.get_10x_lr_params(), "lr": 0.07}], momentum=0.9,
                              weight_decay=5e-4)
        opt_g = torch.optim.Adam(gen.parameters(), lr=2e-4)
        return step_cls(model, gen, crit, crit_g, opt, opt_g, seen, UNSEEN, **kw)

    def time_steps(step, label):
        n0 = L.lib().zs3_launch_count()
        for _ in range(args.warmup):
            step.training_step(image, target, embedding)
        torch.cuda.synchronize()
        n1 = L.lib().zs3_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        n_updates = 0
        for _ in range(args.steps):
            _, _, g_losses = step.training_step(image, target, embedding)
            n_updates += len(g_losses)
        ev1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / args.steps
        ms = ev0.elapsed_time(ev1) / args.steps
        n2 = L.lib().zs3_launch_count()
        return {"impl": label, "ms_per_step": ms, "wall_ms_per_step": wall, "images_per_sec": args.batch / ms * 1e3,
                "generator_updates_per_step": n_updates / args.steps, "native_launches_per_step": (n2 - n1) / args.steps,
                "last_g_loss": g_losses[-1] if g_losses else None}

    res = {"workload": f"ZS3Net step-2 iteration (BASELINE configs[2]), bs={args.batch} {args.size}x{args.size}, "
                       f"{C} classes, unseen {UNSEEN}", "steps": args.steps, "warmup": args.warmup, "runs": []}
    fused = build(ZS3StepFused)
    res["runs"].append(time_steps(fused, "ZS3StepFused (device work list + zs3_gmmn_train_fused)"))
    try:
        gstep = build(ZS3StepFused, graph_features=True)
        res["runs"].append(time_steps(gstep, "ZS3StepFused + feature extraction replayed from a CUDA graph"))
        gstep.profile = {}
        gstep.training_step(image, target, embedding)
        torch.cuda.synchronize()
        res["segments_of_one_graph_step"] = gstep.profile_summary()
        gstep.profile = None
    except Exception as e:  # the graph variant is opt-in; report instead of losing the other numbers
        res["runs"].append({"impl": "ZS3StepFused graph_features", "error": repr(e)[:300]})
    res["runs"].append(time_steps(build(ZS3StepFused, graph_features=True, tensor_core_bulk=False),
                                  "ZS3StepFused + graph, unseen-image features per class on the fp32 SIMT GEMM"))
    if not args.skip_unfused:
        res["runs"].append(time_steps(build(ZS3Step), "ZS3Step (module by module)"))

    # feature extraction alone (the DeepLab forward both variants share)
    model = fused.model.module
    with torch.no_grad():
        for _ in range(2):
            model.forward_before_class_prediction(image)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            feats = model.forward_before_class_prediction(image)
        ev1.record()
        torch.cuda.synchronize()
    res["feature_extraction_ms"] = ev0.elapsed_time(ev1) / args.steps

    # the fused kernel alone on a 48-update work list (sampled rows of real feature maps)
    from zs3_b200 import gmmn_fused as GF
    feats = feats.contiguous().float()
    hw = feats.shape[2] * feats.shape[3]
    items, keep = [], []
    for k in range(48):
        i = k % args.batch
        pix = torch.randint(0, hw, (128,), device=dev, dtype=torch.int32)
        z = torch.rand(128, 300, device=dev)
        keep += [pix, z]
        items.append(GF.pack_item(GF.row_source(emb_table[k % C:k % C + 1], row_stride=0), GF.row_source(z),
                                  GF.row_source(feats[i], pix, row_stride=1, col_stride=hw), 128))
    upd = fused.updater
    for _ in range(2):
        upd.run(items, 300, 300, keepalive=keep)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    reps = 5
    for _ in range(reps):
        losses = upd.run(items, 300, 300, keepalive=keep)
    ev1.record()
    torch.cuda.synchronize()
    per_update_us = ev0.elapsed_time(ev1) * 1e3 / (reps * len(items))
    # algorithmic work of one update (fp32 FMA): fwd 128x600x256 + 128x256x256, MMD 256x256x256 (+ the 128-row
    # gradient 128x256x256), bwd 2 x 128x256x256 + 128x256x600
    mfma = (128 * 600 * 256 * 2 + 128 * 256 * 256 * 4 + 256 * 256 * 256) / 1e6
    res["fused_kernel"] = {"updates_per_launch": len(items), "us_per_update": per_update_us,
                           "gflops_fp32": 2 * mfma / per_update_us * 1e3,
                           "mfma_per_update": mfma, "finite_losses": bool(torch.isfinite(losses).all())}
    line = json.dumps(res)
    print(line)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
