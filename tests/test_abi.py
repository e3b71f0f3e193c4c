"""CPU: the C-ABI library loads and exports every symbol include/zs3b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "zs3b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zs3_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from zs3_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with `python -c 'import __graft_entry__ as g; g.build()'`"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_header_and_reports_errors():
    from zs3_b200 import _lib
    lib = _lib.lib()
    bound = set(lib._zs3_declared)
    declared = set(_declared_symbols())
    assert declared <= bound, sorted(declared - bound)
    assert lib.zs3_abi_version() == 1
    # argument validation happens before any CUDA call: usable without a GPU
    rc = lib.zs3_conv_fprop(None, None)
    assert rc == -1 and b"null args" in lib.zs3_last_error()
    a = _lib.ConvArgs()
    a.num_segments = 9
    rc = lib.zs3_conv_fprop(ctypes.byref(a), None)
    assert rc == -1 and b"num_segments" in lib.zs3_last_error()


def test_binding_struct_sizes_match_the_library():
    """every ctypes mirror has the size the library was compiled with (a stale stub is an OOB read)"""
    from zs3_b200 import _lib
    lib = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "zs3b200.h")).read()
    ids = dict((int(v), k) for k, v in re.findall(r"#define (ZS3_STRUCT_[A-Z0-9_]+) (\d+)", hdr))
    assert sorted(ids) == sorted(_lib.STRUCT_IDS), (ids, _lib.STRUCT_IDS)
    for which, mirror in _lib.STRUCT_IDS.items():
        assert lib.zs3_sizeof(which) == ctypes.sizeof(mirror) > 0, (ids[which], mirror)
    assert lib.zs3_sizeof(999) == 0
    # the stub printed in INTEGRATION.md is the binding's own struct, not a hand copy
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert "from zs3_b200._lib import ConvArgs" in doc and "class ConvArgs(C.Structure)" not in doc


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import importlib
    from zs3_b200 import _lib
    importlib.reload(_lib)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    import pytest
    with pytest.raises(_lib.Zs3NativeError):
        _lib.lib()
    importlib.reload(_lib)


def test_product_code_never_imports_the_oracle():
    bad = []
    for base in ("zs3_b200", "zs3"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if (any(name in txt for name in ("zs3_oracle", "zs3_step2_oracle", "zs3_graph_oracle"))
                            or "import oracle" in txt or "from oracle" in txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_zs3_shim_falls_through_to_a_reference_checkout(tmp_path):
    """With this repo before a reference checkout on sys.path, hot-path modules resolve here and the rest of the
    reference's namespace package (dataloaders, parsing, utils.saver, ...) stays importable."""
    import subprocess
    import sys
    ref = tmp_path / "ref" / "zs3"
    (ref / "utils").mkdir(parents=True)
    (ref / "dataloaders").mkdir()
    (ref / "parsing.py").write_text("WHO = 'reference parsing'\n")
    (ref / "utils" / "lr_scheduler.py").write_text("WHO = 'reference lr_scheduler'\n")
    (ref / "utils" / "loss.py").write_text("WHO = 'reference loss (must be shadowed)'\n")
    (ref / "dataloaders" / "__init__.py").write_text("WHO = 'reference dataloaders'\n")
    code = ("import zs3.parsing, zs3.utils.lr_scheduler, zs3.dataloaders, zs3.utils.loss, zs3.utils.metrics;"
            "print(zs3.parsing.WHO, '|', zs3.utils.lr_scheduler.WHO, '|', zs3.dataloaders.WHO, '|',"
            "hasattr(zs3.utils.loss, 'GMMNLoss'), hasattr(zs3.utils.metrics, 'Evaluator'))")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, str(tmp_path / "ref")]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "reference parsing | reference lr_scheduler | reference dataloaders | True True"


def test_new_entry_points_validate_arguments_without_a_gpu():
    """argument checks of the round-1b entry points happen before any CUDA call (usable on the CPU box)"""
    import ctypes as C
    from zs3_b200 import _lib
    lib = _lib.lib()
    err = lambda: lib.zs3_last_error().decode()  # noqa: E731
    # fused generator update: null args, bad dims, workspace too small
    assert lib.zs3_gmmn_train_fused(None, None) == -1 and "null args" in err()
    a = _lib.GmmnTrainArgs()
    assert lib.zs3_gmmn_train_fused(C.byref(a), None) == -1 and "bad dims" in err()
    need = lib.zs3_gmmn_train_workspace_size(300, 300, 256, 256)
    assert 1_000_000 < need < 4_000_000
    assert lib.zs3_gmmn_train_workspace_size(0, 300, 256, 256) == 0
    buf = (C.c_float * 16)()
    a.embed_dim, a.noise_dim, a.hidden, a.feat, a.nsigma, a.max_rows, a.apply_adam = 300, 300, 256, 256, 6, 128, 0
    for name in ("w1", "b1", "w2", "b2", "losses", "workspace"):
        setattr(a, name, C.addressof(buf))
    for i in range(4):
        a.grad[i] = C.addressof(buf)
    a.workspace_bytes = 64
    assert lib.zs3_gmmn_train_fused(C.byref(a), None) == -1 and "workspace too small" in err()
    a.max_rows = 129
    assert lib.zs3_gmmn_train_fused(C.byref(a), None) == -1 and "max_rows" in err()
    # cluster graph: a map that does not fit in shared memory; null pointers
    c = _lib.ComponentsArgs()
    assert lib.zs3_label_components(C.byref(c), None) == -1 and "null pointer" in err()
    for name in ("labels", "n_nodes", "node_label", "node_seed", "adj"):
        setattr(c, name, C.addressof(buf))
    c.B, c.h, c.w, c.max_nodes = 1, 400, 400, 8
    assert lib.zs3_label_components(C.byref(c), None) == -1 and "shared memory" in err()
    # metrics: more classes than the per-CTA histogram holds; nothing to compute
    p = C.addressof(buf)
    assert lib.zs3_argmax_confusion(p, p, 1, 65, 10, p, p, None) == -1 and "1 <= C <= 64" in err()
    assert lib.zs3_argmax_confusion(p, None, 1, 21, 10, None, None, None) == -1 and "nothing to compute" in err()
    assert lib.zs3_confusion_from_pred(None, p, 10, 21, p, None) == -1


def test_augment_entry_point_validates_arguments_without_a_gpu():
    import ctypes as C
    from zs3_b200 import _lib
    lib = _lib.lib()
    err = lambda: lib.zs3_last_error().decode()  # noqa: E731
    assert lib.zs3_augment_batch(None, None) == -1 and "null args" in err()
    a = _lib.AugmentArgs()
    assert lib.zs3_augment_batch(C.byref(a), None) == -1 and "bad dims" in err()
    assert lib.zs3_augment_workspace_size(0, 10, 10, 10) == 0
    need = lib.zs3_augment_workspace_size(16, 500, 513, 513)
    assert 16 * (500 + 513) * 513 * 3 <= need < 16 * (500 + 513) * 513 * 3 + 16 * (1026 * 20 + 8) * 4 + 64
    buf = (C.c_float * 64)()
    p = C.addressof(buf)
    items = (_lib.AugItem * 1)()
    a.n, a.max_src_h, a.out_w, a.out_h, a.fill_label = 1, 50, 33, 33, 255
    a.items, a.items_host, a.lut, a.out_image, a.out_label, a.workspace = p, items, p, p, p, p
    items[0] = _lib.AugItem(p, p, 40, 50, 0, 4, 50, 0, 0, -1.0)
    assert lib.zs3_augment_batch(C.byref(a), None) == -1 and "8x" in err()
    items[0] = _lib.AugItem(p, None, 40, 50, 0, 40, 50, 0, 0, -1.0)
    assert lib.zs3_augment_batch(C.byref(a), None) == -1 and "label map missing" in err()
    items[0] = _lib.AugItem(p, p, 40, 60, 0, 40, 60, 0, 0, -1.0)
    assert lib.zs3_augment_batch(C.byref(a), None) == -1 and "max_src_h" in err()
    items[0] = _lib.AugItem(p, p, 40, 50, 0, 40, 50, 0, 0, -1.0)
    a.workspace_bytes = 64
    assert lib.zs3_augment_batch(C.byref(a), None) == -1 and "workspace" in err()


def test_wgrad_tile_cost_model_choices_without_a_gpu():
    """host-side tile choice of zs3_conv_wgrad (csrc/conv_igemm.cu choose_wgrad_tile): layer3's 1x1 1024<->256 layers at
    33x33 move to 128 x 128 tiles (9 pixel splits instead of 37), every other shape of the network keeps the largest
    tile (profiles/r02_wgrad_tile.md); ZS3_WGRAD_TILE is not set in the test environment"""
    import ctypes as C
    import os
    from zs3_b200 import _lib
    assert "ZS3_WGRAD_TILE" not in os.environ
    lib = _lib.lib()

    def tile(M, cout, cin, taps):
        mb, cn = C.c_int(0), C.c_int(0)
        assert lib.zs3_debug_wgrad_tile(M, cout, cin, taps, C.byref(mb), C.byref(cn)) == 0
        return mb.value, cn.value

    m33, m65, m129 = 16 * 33 * 33, 16 * 65 * 65, 16 * 129 * 129
    for cout, cin in ((256, 1024), (1024, 256)):
        mb, cn = tile(m33, cout, cin, 1)
        assert 128 * mb * cn < 256 * 256, (cout, cin, mb, cn)     # a smaller tile than the default: fewer pixel splits
    assert tile(m33, 256, 256, 9) == (2, 256)                     # 3x3 256->256 at 33x33: measured slower on small tiles
    assert tile(m129, 256, 256, 9) == (2, 256) and tile(m129, 256, 320, 9) == (2, 256)     # decoder
    assert tile(m33, 256, 2048, 9) == (2, 256) and tile(m33, 512, 512, 9) == (2, 256)      # ASPP, layer4
    assert tile(m129, 64, 64, 9) == (1, 64) and tile(m65, 128, 128, 9) == (1, 128)         # layer1 / layer2 3x3
    assert lib.zs3_debug_wgrad_tile(0, 256, 256, 1, None, None) == -1


def test_reference_scripts_import_surface_resolves_through_the_shim():
    """Every name the reference's trainer / evaluation scripts import from the modules this repository provides
    (zs3.modeling.*, zs3.utils.loss, zs3.utils.metrics) exists under the same import path here (eval_pascal.py
    imports Evaluator_seen_unseen).  Needs the reference checkout (build container only)."""
    import ast
    import glob
    import importlib
    import os
    import pytest
    ref = "/root/reference/zs3"
    if not os.path.isdir(ref):
        pytest.skip("no reference checkout on this machine")
    provided = ("zs3.modeling.deeplab", "zs3.modeling.gmmn", "zs3.modeling.aspp", "zs3.modeling.decoder",
                "zs3.modeling.backbone", "zs3.modeling.sync_batchnorm.replicate", "zs3.modeling.sync_batchnorm.batchnorm",
                "zs3.modeling.sync_batchnorm", "zs3.utils.loss", "zs3.utils.metrics")
    wanted = {}
    for path in glob.glob(os.path.join(ref, "*.py")):
        for node in ast.walk(ast.parse(open(path).read())):
            if isinstance(node, ast.ImportFrom) and node.module in provided:
                for a in node.names:
                    wanted.setdefault(node.module, set()).add(a.name)
    assert wanted.get("zs3.utils.metrics", set()) >= {"Evaluator"} and "zs3.modeling.deeplab" in wanted
    missing = []
    for mod, names in sorted(wanted.items()):
        m = importlib.import_module(mod)
        assert "/root/reference" not in (getattr(m, "__file__", "") or ""), f"{mod} resolved to the reference checkout"
        missing += [f"{mod}.{n}" for n in sorted(names) if not hasattr(m, n)]
    assert not missing, missing
