"""Device-side validation metrics (csrc/metrics.cu, zs3.utils.metrics.Evaluator) vs the reference Evaluator's golden
vectors and the numpy oracle at the full validation size (16 x 21 x 513 x 513 logits); bit-exact integer counts."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

pytestmark = pytest.mark.gpu
SEEN, UNSEEN = [c for c in range(21) if c not in (10, 14)], [10, 14]


def test_evaluator_matches_reference_golden():
    from zs3.utils.metrics import Evaluator
    gold = np.load(os.path.join(HERE, "golden", "metrics.npz"))
    ev = Evaluator(21, SEEN, UNSEEN)
    ev.add_batch(gold["gt"][:2], gold["pred"][:2])                                    # numpy, as the trainer passes them
    ev.add_batch(torch.from_numpy(gold["gt"][2:]).cuda(), torch.from_numpy(gold["pred"][2:]).cuda())
    assert np.array_equal(ev.confusion_matrix, gold["confusion"])
    acc, acc_c, miou, fw = (ev.Pixel_Accuracy(), ev.Pixel_Accuracy_Class(), ev.Mean_Intersection_over_Union(),
                            ev.Frequency_Weighted_Intersection_over_Union())
    assert np.allclose(np.array(acc), gold["pixel_acc"], rtol=1e-12)
    assert np.allclose(np.array([acc_c[0], acc_c[2], acc_c[3]]), gold["class_acc"], rtol=1e-12)
    assert np.allclose(acc_c[1], gold["class_acc_by_class"], rtol=1e-12, equal_nan=True)
    assert np.allclose(np.array([miou[0], miou[2], miou[3]]), gold["miou"], rtol=1e-12)
    assert np.allclose(miou[1], gold["miou_by_class"], rtol=1e-12, equal_nan=True)
    assert np.allclose(np.array(fw), gold["fwiou"], rtol=1e-12)
    ev.reset()
    assert ev.confusion_matrix.sum() == 0
    plain = Evaluator(21)
    plain.add_batch(gold["gt"], gold["pred"])
    assert np.isclose(plain.Pixel_Accuracy(), gold["pixel_acc"][0]) and len(plain.Mean_Intersection_over_Union()) == 2


def test_argmax_confusion_full_size_vs_numpy():
    """the validation batch of zs3/train_pascal_GMMN.py:358-375 at 16 x 21 x 513 x 513: predictions and counts
    bit-exact against np.argmax / the oracle; ties resolve to the first maximum; sum(conf) = labelled pixels"""
    import zs3_oracle as O
    from zs3.utils.metrics import Evaluator
    g = torch.Generator(device="cuda").manual_seed(3)
    B, Cn, H = 16, 21, 513
    logits = torch.randn(B, Cn, H, H, generator=g, device="cuda")
    logits[:, 9] = logits[:, 4]                                                       # exact ties everywhere
    target = torch.randint(0, Cn, (B, H, H), generator=g, device="cuda").float()
    target[torch.rand(B, H, H, generator=g, device="cuda") < 0.02] = 255
    ev = Evaluator(Cn, SEEN, UNSEEN)
    pred = ev.add_batch_logits(target, logits, want_pred=True)
    ev.add_batch_logits(target[:3], logits[:3])
    torch.cuda.synchronize()
    ref_pred = np.argmax(logits.cpu().numpy(), axis=1)
    assert np.array_equal(pred.cpu().numpy(), ref_pred)
    ref_cm = O.confusion_matrix(target.cpu().numpy(), ref_pred, Cn) + O.confusion_matrix(target[:3].cpu().numpy(),
                                                                                         ref_pred[:3], Cn)
    assert np.array_equal(ev.confusion_matrix, ref_cm.astype(np.float64))
    assert int(ev.confusion_matrix.sum()) == int((target != 255).sum().item() + (target[:3] != 255).sum().item())
    sc = O.evaluator_scores(ref_cm, SEEN, UNSEEN)
    assert np.allclose(ev.Mean_Intersection_over_Union()[0], sc["all"][2], rtol=1e-12)


def test_evaluator_on_deeplab_logits_and_errors():
    from zs3.modeling.deeplab import DeepLab
    from zs3.utils.metrics import Evaluator
    model = DeepLab(num_classes=21, pretrained=False).cuda().eval()
    x = torch.randn(2, 3, 65, 65, device="cuda")
    target = torch.randint(0, 21, (2, 65, 65), device="cuda").float()
    with torch.no_grad():
        out = model(x)
    ev = Evaluator(21)
    pred = ev.add_batch_logits(target, out, want_pred=True)
    assert torch.equal(pred.long(), out.argmax(1))
    assert int(ev.confusion_matrix.sum()) == target.numel()
    with pytest.raises(ValueError):
        Evaluator(65)
    with pytest.raises(ValueError):
        ev.add_batch_logits(target, out[:, :20])
    with pytest.raises(RuntimeError):
        ev.add_batch_logits(target.cpu(), out.cpu())
