"""CPU-side checks of the C-ABI library and the host mirror (no compute calls, no GPU)."""
import ctypes
import os
import re

import pytest
import torch

import enf_pde_b200 as E
from enf_pde_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = E.load_library()
    header = open(os.path.join(ROOT, "include", "enf_b200.h")).read()
    declared = set(re.findall(r"\b(enf_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.enf_abi_version() == _lib.ABI_VERSION == 2


def test_struct_layouts_match_header():
    header = open(os.path.join(ROOT, "include", "enf_b200.h")).read()
    body = header.split("typedef struct EnfWeights {")[1].split("} EnfWeights;")[0]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\*\s*([a-z0-9_]+)", body)
    assert tuple(names) == _lib.LEAVES
    assert len(_lib.LEAVES) == 46 and set(_lib.LEAF_PATHS) == set(_lib.LEAVES)
    assert ctypes.sizeof(_lib.EnfDesc) == 64 and ctypes.sizeof(_lib.EnfWeights) == 46 * 8


def test_invariant_dims_match_reference_classes():
    lib = E.load_library()
    import types
    for t, kind in _lib.INVARIANT_KINDS.items():
        dx = 3 if t.startswith("ball") else 2
        inv = E.get_ca_invariant(types.SimpleNamespace(invariant_type=t, num_in=dx))
        assert lib.enf_invariant_dim(kind, dx) == inv.dim
        assert lib.enf_pose_dim(kind, dx) == inv.pose_dim
    assert lib.enf_invariant_dim(99, 2) < 0


def test_bad_descriptions_are_rejected_without_a_gpu():
    lib = E.load_library()
    ok = dict(B=2, C=10, Z=4, d=32, H=2, L=8, O=1, Dx=2, invariant_kind=3, use_window=1, precision=0, flags=0)
    assert lib.enf_xattn_workspace_bytes(ctypes.byref(_lib.EnfDesc(**ok))) > 0
    for bad in (dict(d=48), dict(H=5), dict(B=0), dict(invariant_kind=42), dict(Dx=3), dict(precision=7), dict(flags=32), dict(flags=_lib.FLAG_FROZEN_RELU, precision=1),
                dict(flags=_lib.FLAG_OUT_BF16), dict(chunk_fields=-1), dict(B=400, Z=64, H=4)):
        d = _lib.EnfDesc(**{**ok, **bad})
        assert lib.enf_xattn_workspace_bytes(ctypes.byref(d)) == 0, bad
        assert lib.enf_last_error() != b""


def test_limits_and_dispatch_are_reported():
    """B*Z*H <= 65535 per call (the stage kernels' grid z-dimension); enf_xattn_dispatch names the kernels a description gets."""
    lib = E.load_library()
    ok = dict(B=32, C=4096, Z=64, d=128, H=2, L=16, O=1, Dx=2, invariant_kind=3, use_window=1, precision=1, flags=0)
    assert _lib.dispatch(_lib.EnfDesc(**ok)) == (True, True)
    assert _lib.dispatch(_lib.EnfDesc(**{**ok, "precision": 0})) == (False, False)
    assert _lib.dispatch(_lib.EnfDesc(**{**ok, "d": 32, "H": 3})) == (True, True)            # config 5 (ihc): tcgen05 kernels
    assert _lib.dispatch(_lib.EnfDesc(**{**ok, "d": 16, "H": 2})) == (False, False)          # d = 16: fp32 kernels in either mode
    assert _lib.dispatch(_lib.EnfDesc(**{**ok, "flags": _lib.FLAG_FORWARD_ONLY})) == (True, False)
    big = _lib.EnfDesc(**{**ok, "B": 512})                                                      # 512 * 64 * 2 = 65536
    assert lib.enf_xattn_workspace_bytes(ctypes.byref(big)) == 0 and b"65535" in lib.enf_last_error()
    assert lib.enf_xattn_dispatch(ctypes.byref(big), None, None) == -2
    edge = _lib.EnfDesc(**{**ok, "B": 511, "C": 128, "flags": _lib.FLAG_FORWARD_ONLY})
    assert lib.enf_xattn_workspace_bytes(ctypes.byref(edge)) > 0


def test_recompute_mode_bounds_the_workspace():
    """ENF_FLAG_RECOMPUTE: nothing O(B*C*Z) is kept between forward and backward; the workspace scales with chunk_fields and
    enf_xattn_chunk_for_cap inverts that.  ns64 (BASELINE config 2) trains within 2.75 GiB (8.9 GiB with the stash)."""
    lib = E.load_library()
    kw = dict(B=32, C=4096, Z=64, d=128, H=2, L=16, O=1, Dx=2, invariant_kind=3, use_window=1, precision=1)
    size = lambda **o: lib.enf_xattn_workspace_bytes(ctypes.byref(_lib.EnfDesc(**{**kw, **o})))
    stash = size(flags=0)
    sizes = [size(flags=_lib.FLAG_RECOMPUTE, chunk_fields=n) for n in (1, 2, 4, 8, 32)]
    assert all(a < b for a, b in zip(sizes, sizes[1:])) and sizes[-1] == stash
    assert size(flags=_lib.FLAG_RECOMPUTE) == sizes[2]                                          # default chunk: 4 fields
    assert sizes[0] < 2.75 * 2 ** 30 and sizes[0] < 0.32 * stash, sizes[0]
    d = _lib.EnfDesc(**kw, flags=0)
    for cap_gib, in (3.0,), (4.0,), (100.0,):
        n = lib.enf_xattn_chunk_for_cap(ctypes.byref(d), int(cap_gib * 2 ** 30))
        assert 1 <= n <= 32 and size(flags=_lib.FLAG_RECOMPUTE, chunk_fields=n) <= cap_gib * 2 ** 30
        assert n == 32 or size(flags=_lib.FLAG_RECOMPUTE, chunk_fields=n + 1) > cap_gib * 2 ** 30
    assert lib.enf_xattn_chunk_for_cap(ctypes.byref(d), 1 << 20) == 0
    # the fp32 kernels always recompute: the flag changes nothing there
    assert size(precision=0, flags=_lib.FLAG_RECOMPUTE) == size(precision=0, flags=0)


def test_jax_binding_module_is_importable_and_refuses_without_jax():
    """the jax.ffi glue is shipped as source; without JAX it must import and fail with a clear message, not crash."""
    import enf_pde_b200.jax_binding as J
    if J.HAVE_JAX:
        pytest.skip("JAX present: the binding needs the shim library, built separately")
    with pytest.raises(ImportError, match="jax"):
        J.make_enf_apply(128, 2, 1, 16, "rel_pos_periodic", 2)
    src = open(os.path.join(os.path.dirname(J.__file__), "csrc", "enf_xla_ffi.cc")).read()
    for sym in ("EnfXattnFwd", "EnfXattnBwd", "enf_xattn_fwd(", "enf_xattn_bwd(", "XLA_FFI_DEFINE_HANDLER_SYMBOL"):
        assert sym in src


def test_xla_ffi_shim_type_checks():
    """csrc/enf_xla_ffi.cc cannot be built against jaxlib in this image; it is at least type-checked on every run against a
    minimal stand-in for xla/ffi/api/ffi.h (tests/mock_xla) that declares the part of the typed FFI API the shim uses."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    src = os.path.join(ROOT, "enf_pde_b200", "csrc", "enf_xla_ffi.cc")
    cmd = [gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "mock_xla"),
           "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "include"), src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(src).read()
    assert "const_cast<uint8_t*>" not in text            # the backward writes through an aliased RESULT, not an input operand
    assert "workspace_out" in text and "chunk_fields" in text


def test_forward_only_workspace_is_smaller():
    """ENF_FLAG_FORWARD_ONLY (validation roll-outs): no backward state in the workspace."""
    lib = E.load_library()
    for prec in (0, 1):
        kw = dict(B=32, C=4096, Z=64, d=128, H=2, L=16, O=1, Dx=2, invariant_kind=3, use_window=1, precision=prec)
        train = lib.enf_xattn_workspace_bytes(ctypes.byref(_lib.EnfDesc(**kw, flags=0)))
        infer = lib.enf_xattn_workspace_bytes(ctypes.byref(_lib.EnfDesc(**kw, flags=_lib.FLAG_FORWARD_ONLY)))
        assert 0 < infer < 0.6 * train, (prec, train, infer)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_compute_fails_loudly_without_a_device():
    """no CPU fallback: the public API refuses CPU tensors, and the C ABI reports ENF_ERR_NO_DEVICE."""
    import types
    inv = E.get_ca_invariant(types.SimpleNamespace(invariant_type="rel_pos_periodic", num_in=2))
    nef = E.EquivariantCrossAttentionNeF(32, 2, 0, 1, 8, inv)
    p, a, s = E.init_latents(inv, 1, 4, 8)
    x = torch.zeros(1, 5, 2)
    params = nef.init(0, x, p, a, s)
    with pytest.raises(RuntimeError):
        nef.apply(params, x, p, a, s)
    lib = E.load_library()
    d = _lib.EnfDesc(B=1, C=5, Z=4, d=32, H=2, L=8, O=1, Dx=2, invariant_kind=3, use_window=1, precision=0, flags=0)
    w = _lib.EnfWeights(**{n: 16 for n in _lib.LEAVES})
    rc = lib.enf_xattn_fwd(ctypes.byref(d), ctypes.byref(w), 16, 10, 16, 16, 16, 16, 256, 1 << 40, None)
    assert rc == -6 and b"no CUDA device" in lib.enf_last_error()


def test_init_tree_matches_reference_names_and_shapes():
    import types
    from helpers import load_golden
    from oracle import enf_ref as R
    cfg, params, _, rec = load_golden("ball")
    inv = E.get_ca_invariant(types.SimpleNamespace(invariant_type="ball", num_in=3))
    nef = E.EquivariantCrossAttentionNeF(cfg.num_hidden, cfg.num_heads, 0, cfg.num_out, cfg.latent_dim, inv,
                                         embedding_freq_multiplier=cfg.embedding_freq_multiplier)
    ours = R.tree_flatten(nef.init(0, rec["x"], rec["p"], rec["a"], rec["sigma"])["params"])
    theirs = R.tree_flatten(params["params"])
    assert sorted(ours) == sorted(theirs)
    for k in ours:
        assert tuple(ours[k].shape) == tuple(theirs[k].shape), k
    # round trip through the flat leaf order of the C ABI
    leaves = E.params_to_leaves({"params": R.tree_unflatten(ours)})
    back = R.tree_flatten(E.leaves_to_params(leaves)["params"])
    assert all(back[k] is ours[k] for k in ours)


def test_ode_abi_exports_and_layouts():
    """include/enf_ode_b200.h (latent ODE model): every declared entry point is exported, the ctypes mirrors have the
    header's layout, bad descriptions are rejected without a GPU."""
    from enf_pde_b200 import ode
    lib = ode._load()
    header = open(os.path.join(ROOT, "include", "enf_ode_b200.h")).read()
    declared = set(re.findall(r"\b(enf_(?:mlp)?ode_[a-z_0-9]+)\s*\(", header))
    assert declared == set(ode.EXPORTS), declared ^ set(ode.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    body = re.sub(r"/\*.*?\*/", "", header.split("typedef struct EnfOdeLayer {")[1].split("} EnfOdeLayer;")[0], flags=re.S)
    assert tuple(re.findall(r"\*\s*([a-z0-9_]+)", body)) == ode._LAYER_LEAVES
    assert int(re.search(r"#define ENF_ODE_MAX_LAYERS (\d+)", header).group(1)) == ode.MAX_LAYERS
    assert ctypes.sizeof(ode.EnfOdeDesc) == 64 and ctypes.sizeof(ode.EnfOdeLayer) == 64
    assert ctypes.sizeof(ode.EnfOdeWeights) == 8 * (5 + 8 * ode.MAX_LAYERS + 3)
    ok = dict(B=2, Z=8, L=16, hidden=128, basis=64, layers=3, widen=2, degree=3, Dx=2, invariant_kind=3)
    assert lib.enf_ode_workspace_bytes(ctypes.byref(ode.EnfOdeDesc(**ok))) > 0
    for bad in (dict(layers=0), dict(layers=9), dict(degree=6), dict(B=0), dict(invariant_kind=42), dict(Dx=3), dict(hidden=0)):
        assert lib.enf_ode_workspace_bytes(ctypes.byref(ode.EnfOdeDesc(**{**ok, **bad}))) == 0, bad
        assert lib.enf_last_error()
    assert ctypes.sizeof(ode.EnfMlpOdeDesc) == 32 and ctypes.sizeof(ode.EnfMlpOdeWeights) == 16 * 8
    mok = dict(B=2, Z=8, P=2, L=16, hidden=64)
    assert lib.enf_mlpode_workspace_bytes(ctypes.byref(ode.EnfMlpOdeDesc(**mok))) > 0
    assert lib.enf_mlpode_workspace_bytes(ctypes.byref(ode.EnfMlpOdeDesc(**{**mok, "P": 3}))) == 0      # 2-component pose derivative
    # compute entry points refuse NULL arguments before touching the device
    d = ode.EnfOdeDesc(**ok)
    assert lib.enf_ode_fwd(ctypes.byref(d), None, None, None, None, None, None, 0, None) == -3


def test_ode_host_mirror_matches_reference_structure():
    """PonitaODEGen.init builds the reference's parameter tree (names / shapes from tests/golden/ode_ponita.npz, written by the
    reference's own module); get_sa_invariant mirrors invariant/__init__.py:13-45."""
    import types
    from helpers import load_ode_golden
    from oracle import enf_ref as R
    cfg, params, _, rec = load_ode_golden("ponita")
    inv = E.get_sa_invariant(types.SimpleNamespace(invariant_type="ponita", num_in=2))
    assert (inv.dim, inv.num_x_ori_dims, inv.pose_dim) == (3, 1, 3)
    assert E.get_sa_invariant(types.SimpleNamespace(invariant_type="rel_pos_periodic", num_in=2)).dim == 4
    m = E.PonitaODEGen(cfg.num_hidden, cfg.num_layers, cfg.latent_dim, 1, inv, cfg.basis_dim, cfg.degree, cfg.widening_factor)
    ours = R.tree_flatten(m.init(0, (rec["p"], rec["a"], rec["sigma"]), device="cpu")["params"])
    theirs = R.tree_flatten(params)
    assert sorted(ours) == sorted(theirs)
    for k in ours:
        assert tuple(ours[k].shape) == tuple(theirs[k].shape), k
    assert float(ours["ponita/readout_vec_rel/kernel"].abs().max()) < 1e-2        # variance_scaling(1e-6)
    with pytest.raises(NotImplementedError):
        E.PonitaODEGen(16, 2, 4, 1, inv, 8, 3, 2, kernel_size=0.5)


def test_self_block_flags_are_validated():
    """ENF_FLAG_SELF_BLOCK / ENF_FLAG_NO_STEM (latent self-attention steps): shape contract checked without a GPU."""
    lib = E.load_library()
    ok = dict(B=2, C=8, Z=8, d=32, H=2, L=6, O=32, Dx=2, invariant_kind=4, use_window=1, precision=1, flags=_lib.FLAG_SELF_BLOCK)
    size = lambda **o: lib.enf_xattn_workspace_bytes(ctypes.byref(_lib.EnfDesc(**{**ok, **o})))
    assert size() > 0
    assert _lib.dispatch(_lib.EnfDesc(**ok)) == (False, False)                       # always the fp32 kernels
    assert size(C=9) == 0 and b"C == Z" in lib.enf_last_error()                      # the queries are the latents
    assert size(O=1) == 0
    assert size(flags=_lib.FLAG_SELF_BLOCK | _lib.FLAG_NO_STEM) == 0 and b"L must equal d" in lib.enf_last_error()
    assert size(flags=_lib.FLAG_SELF_BLOCK | _lib.FLAG_NO_STEM, L=32) > 0
    assert size(flags=_lib.FLAG_SELF_BLOCK | _lib.FLAG_RECOMPUTE) == 0
    assert size(flags=_lib.FLAG_NO_STEM, C=100, O=1, L=32) > 0                      # the decode call after self-attention steps
    # leaves a self-attention step does not use may be NULL; the ones it uses may not
    paths = _lib.leaf_paths("self_attention_blocks_1", with_stem=False, with_mlp=False)
    assert paths["stem_w"] is None and paths["m2_w"] is None and paths["wo"] == "self_attention_blocks_1/attn/out_proj/kernel"
