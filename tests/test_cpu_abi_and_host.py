"""CPU-only checks: the C-ABI library builds/loads and exports every symbol include/marf_b200.h declares;
host logic (options, sharding, coefficients); no compute calls (there is no GPU here and no CPU path)."""
import ctypes as C
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from marf_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from marf_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "marf_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(marf_\w+)\s*\(", hdr, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.marf_abi_version() == _lib.MARF_ABI_VERSION


def test_struct_layouts_match_header(lib):
    """field order/count of the ctypes mirrors vs the C structs (names parsed from the header)."""
    from marf_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "marf_b200.h")).read()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            decl = re.sub(r"^(const\s+)?(float|double|int32_t|uint32_t|int64_t)\s*(\*\s*const\s*\*|\*)?\s*", "", stmt)
            for part in decl.split(","):
                names.append(re.sub(r"\[.*\]", "", part).strip())
        return names
    assert fields("marf_config") == [f[0] for f in _lib.MarfConfig._fields_]
    assert fields("marf_step_io") == [f[0] for f in _lib.MarfStepIO._fields_]
    assert fields("marf_render_io") == [f[0] for f in _lib.MarfRenderIO._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from marf_b200 import _lib
    cfg = _lib.MarfConfig()
    cfg.abi_version = _lib.MARF_ABI_VERSION
    h = C.c_void_p()
    rc = lib.marf_create(C.byref(cfg), C.byref(h))
    assert rc == -3 and b"no CPU path" in lib.marf_last_error(None)
    from marf_b200.engine import PlanarEngine
    with pytest.raises(RuntimeError):
        PlanarEngine(H=8, W=8, patch_H=4, patch_W=4, batch_size=1, layers=[3])


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from marf_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "marf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "planar_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f


def test_argument_parser_semantics():
    from marf_b200 import options
    o = options.parse_arguments(["--model=planar", "--yaml=planar", "--barf_c2f=[0,0.4]", "--arch.posenc!", "--seed=3",
                                 "--load=", "--use_implicit_mask"])
    assert o.model == "planar" and o.barf_c2f == [0, 0.4] and o.arch.posenc is False and o.seed == 3
    assert o.load is None and o.use_implicit_mask is True
    with pytest.raises(ValueError):
        options.parse_arguments(["--a=1", "--a=2"])
    with pytest.raises(ValueError):
        options.parse_arguments(["a=1"])


def test_yaml_defaults_match_reference_planar_yaml():
    from marf_b200 import options
    opt = options.load_options("options/planar.yaml")
    assert (opt.H, opt.W, opt.patch_H, opt.patch_W, opt.batch_size, opt.max_iter) == (360, 480, 180, 240, 5, 3000)
    assert opt.arch.layers == [None, 256, 256, 256, 256, 3] and opt.arch.skip == [] and opt.arch.posenc.L_2D == 8
    assert opt.barf_c2f is None and opt.use_masks is True and opt.use_implicit_mask is False and opt.use_edges is True
    assert opt.loss_weight == dict(render=0, rgb=0, edge=0, mask=0)
    assert (opt.optim.lr, opt.optim.lr_warp, opt.optim.lr_mask, opt.optim.algo) == (1e-3, 1e-3, 1e-3, "Adam")
    assert opt.warp.fix_first is True and opt.warp.dof == 8 and opt.N_vocab == 1500
    # unknown keys are refused when stdin is not a TTY (the reference prompts)
    with pytest.raises(KeyError):
        options.override_options(opt, options.parse_arguments(["--not_a_key=1"]), key_stack=[], safe_check=True)


def test_override_and_inheritance(tmp_path, monkeypatch):
    from marf_b200 import options
    (tmp_path / "options").mkdir()
    (tmp_path / "options" / "base.yaml").write_text("a: 1\nb:\n    c: 2\n    d: 3\n")
    (tmp_path / "options" / "child.yaml").write_text("_parent_: options/base.yaml\nb:\n    c: 5\ne: 6\n")
    monkeypatch.chdir(tmp_path)
    opt = options.load_options("options/child.yaml")
    assert opt.a == 1 and opt.b.c == 5 and opt.b.d == 3 and opt.e == 6


def test_shard_plan():
    from marf_b200.engine import shard_plan
    assert shard_plan(5, 180, 0, 1) == (5, 0, 180, 0)
    assert [shard_plan(64, 1024, r, 8) for r in (0, 7)] == [(8, 0, 1024, 0), (8, 56, 1024, 0)]
    parts = [shard_plan(5, 180, r, 8) for r in range(8)]            # rows split when patches do not divide
    assert all(p[0] == 5 and p[1] == 0 for p in parts)
    assert sum(p[2] for p in parts) == 180 and parts[0][3] == 0
    for a, b in zip(parts[:-1], parts[1:]):
        assert a[3] + a[2] == b[3]
    # the edge term needs whole patches: dealt out unevenly, possibly none for the last ranks
    assert [shard_plan(5, 180, r, 2, whole_patches=True) for r in range(2)] == [(3, 0, 180, 0), (2, 3, 180, 0)]
    parts = [shard_plan(5, 180, r, 8, whole_patches=True) for r in range(8)]
    assert [p[0] for p in parts] == [1, 1, 1, 1, 1, 0, 0, 0] and [p[1] for p in parts[:5]] == [0, 1, 2, 3, 4]
    assert shard_plan(64, 1024, 3, 8, whole_patches=True) == (8, 24, 1024, 0)


def test_loss_coefficients_match_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import planar_oracle as po
    from marf_b200.attrdict import AttrDict
    from marf_b200.planar import Graph
    for lw in (dict(render=0, rgb=0, edge=0, mask=0), dict(render=0, rgb=-1, edge=0.5, mask=None), dict(render=1, rgb=None, edge=0, mask=0)):
        for use_edges in (False, True):
            cfg = po.PlanarConfig(use_edges=use_edges, loss_weight=dict(lw), max_iter=200)
            fake = AttrDict(opt=AttrDict(use_edges=use_edges, alpha_initial=0.0, alpha_final=1.0, loss_weight=AttrDict(lw)),
                            it=60, max_iter=200)
            got = Graph.loss_coefficients(fake)
            assert got[:3] == pytest.approx(po.loss_coefficients(cfg, 60))
            assert got[3] == pytest.approx(po.edge_alpha(cfg, 60))


def test_attrdict():
    from marf_b200.attrdict import AttrDict
    d = AttrDict(a=1, b=dict(c=[dict(x=1)], d=None))
    assert d.b.c[0].x == 1 and d["b"]["d"] is None
    d.e = dict(f=2)
    assert d.e.f == 2 and d.to_dict() == {"a": 1, "b": {"c": [{"x": 1}], "d": None}, "e": {"f": 2}}
    assert d.pop("a") == 1 and "a" not in d
    with pytest.raises(AttributeError):
        d.missing


def test_no_predicated_tensor_core_mma_in_sass():
    """ptxas 12.9 if-converts a branch around tcgen05.mma into a predicated UTCHMMA whose descriptor moves (R2UR) can end
    up guarded by an unrelated predicate: the MMA then runs with stale descriptors (seen in k_tc_chain).  Every tcgen05.mma
    of the library must therefore be issued unconditionally; this scans the built SASS for a predicated one."""
    import shutil
    import subprocess
    from marf_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    lib = build.build()
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    lines = [l for l in sass.splitlines() if "UTCHMMA" in l]
    assert len(lines) > 50, "tensor-core MMAs missing from the library?"
    bad = [l.strip() for l in lines if "@" in l.split("UTCHMMA")[0]]
    assert not bad, bad[:4]


def test_fp32_mode_kernels_are_tensor_core_kernels_in_sass():
    """precision=fp32 runs its wide layers as 3xTF32 GEMMs (csrc/tc_tf32.cuh): the built k_tf32x3 kernels must hold CTA-pair
    tensor-core MMAs, TMA loads / stores and TMEM loads, and must address shared memory with STS / LDS (through the rounded-up
    dynamic-SMEM pointer the compiler once emitted generic ST.E / LD.E for every operand plane: 2-3x the loaders' issue time)."""
    import collections
    import re
    import shutil
    import subprocess
    from marf_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([cuobjdump, "-sass", build.build()], capture_output=True, text=True, check=True).stdout
    ops, fn, n_fn = collections.Counter(), None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            n_fn += "k_tf32x3" in fn
            continue
        if fn and "k_tf32x3" in fn:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][\w.]*)", line)
            if m:
                ops[m.group(1).split(".")[0] + ("." + m.group(1).split(".")[1] if m.group(1).startswith(("ST.", "LD.")) else "")] += 1
    assert n_fn == 6, n_fn                      # NT x {bias, bias+relu, plain, relu-mask, relu-bits} + TN
    for need in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STS", "LDS"):
        assert ops[need] > 0, (need, dict(ops))
    assert ops["ST.E"] == 0, "generic stores in the 3xTF32 kernels (shared-memory planes must be written with st.shared)"


def test_bench_roofline_arithmetic():
    """bench.roofline_lines: algorithmic FLOP / bytes per kernel class, the dominant kernel against the TENSOR roofline
     with the HBM view and the traffic ratio beside it."""
    import bench
    wl = bench.workload("config2", 1)
    pk = dict(bf16_burst=1600.0, bf16_sustained=1400.0, hbm=6500.0, source="test")
    kern = {"k_tc_bwd": dict(us_per_launch=400.0, launches_per_step=1.0),
            "k_tc_chain<fwd>": dict(us_per_launch=220.0, launches_per_step=1.0),
            "k_tc_gemm<64,warp_grad>": dict(us_per_launch=38.0, launches_per_step=1.0)}
    out = bench.roofline_lines(wl, "config2", "bf16", kern, pk, 216000, 650.0, 1)
    rows = 216064                                        # 216000 padded to 128
    r = out["roofline"]
    assert r["kernel"] == "k_tc_bwd" and r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["peak"] == 1600.0
    # algorithmic FLOP of the three kernel classes add up to SURVEY 8a's 2,853,888 per pixel-sample (image MLP + mask head)
    flop, byt = bench.kernel_tables(wl)
    assert flop["k_tc_chain<fwd>"] + flop["k_tc_bwd"] + flop["k_tc_gemm<64,warp_grad>"] == 2_853_888
    assert flop["k_tc_bwd"] == flop["k_tc_dw"] + flop["k_tc_chain<dx>"]
    flop1, _ = bench.kernel_tables(bench.workload("config4", 1))
    assert flop1["k_tc_chain<fwd>"] + flop1["k_tc_bwd"] + flop1["k_tc_gemm<64,warp_grad>"] == 1_236_480
    assert abs(r["achieved"] - flop["k_tc_bwd"] * rows / 400e-6 / 1e12) < 1e-9
    assert abs(r["frac"] - r["achieved"] / 1600.0) < 1e-12 and abs(r["frac_of_sustained"] - r["achieved"] / 1400.0) < 1e-12
    # two networks: every X_l once (4 x 512 + 128 B), dlogits tiles, mask bits; dY_0 of the image chain out
    assert r["algorithmic_bytes"] == (2 * (4 * 512 + 128 + 32 + 4 * 32) + 512) * rows
    assert abs(r["hbm_view"]["achieved"] - r["algorithmic_bytes"] / 400e-6 / 1e9) < 1e-6
    by = {k["kernel"]: k for k in out["kernels"]}
    assert [k["kernel"] for k in out["kernels"]][0] == "k_tc_bwd"          # sorted by time in the step
    fwd = by["k_tc_chain<fwd>"]
    assert abs(fwd["hbm_view"]["achieved"] - 2 * (128 + 4 * 512 + 4 * 32 + 16) * rows / 220e-6 / 1e9) < 1e-6
    assert out["step_roofline"]["frac"] == 650.0 / 1600.0 and out["step_roofline"]["frac_of_sustained"] == 650.0 / 1400.0
    # traffic comes from the committed ncu capture when it holds this workload
    assert r["traffic"] is None or r["traffic"] > 1e8
    # fp32 / no per-kernel pass: falls back to the whole-step figure
    out2 = bench.roofline_lines(wl, "config2", "bf16", {}, pk, 216000, 650.0, 1)
    assert out2["roofline"]["bound"] == "tensor" and out2["roofline"]["traffic"] is None
