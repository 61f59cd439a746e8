"""CPU tests: host-side mirror of the reference interface, the C-ABI symbol table and (when the reference tree is
mounted) drop-in equivalences against the unmodified reference classes.  No CUDA compute calls."""
import ctypes
import os
import sys
import re
import tempfile

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, params_of

import cmrl_b200 as cm
from oracle import refshim

SILU = {"_target_": "torch.nn.SiLU"}


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cmrl_b200.h")).read()
    declared = set(re.findall(r"\b(cmrl_[a-z0-9_]+)\s*\(", header))
    lib = cm._lib.load()
    assert lib.cmrl_abi_version() == 1
    for name in declared:
        assert hasattr(lib, name), "libcmrl_b200.so does not export %s" % name
    bound = set(cm._lib.PROTOTYPES) | {"cmrl_abi_version", "cmrl_last_error", "cmrl_launch_count", "cmrl_train_fused_sizes"}
    assert declared == bound, declared ^ bound
    # argument validation happens before any CUDA call
    rc = lib.cmrl_ens_linear_fwd(None, 0, 0, None, None, None, None, 1, 1, 1, 1, 1, None, 0, 0, 0, None, 0, 0, None)
    assert rc < 0 and b"null pointer" in lib.cmrl_last_error()


def _header_declarations():
    """[(name, [parameter categories])] of every function include/cmrl_b200.h declares (comments stripped)."""
    text = open(os.path.join(ROOT, "include", "cmrl_b200.h")).read()
    text = re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", text, flags=re.S))
    scalar = {"int": "int", "long long": "ll", "unsigned long long": "ull", "float": "float"}
    out = []
    for name, params in re.findall(r"\b(?:int|long long|const char\*)\s+(cmrl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        cats = []
        for p_ in params.split(","):
            p_ = " ".join(p_.split())
            if p_ == "void":
                continue
            if "*" in p_:
                cats.append("ptr")
            else:
                cats.append(scalar[re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*$", "", p_).replace("const ", "").strip()])
        out.append((name, cats))
    return out


def test_ctypes_prototypes_match_the_header_parameter_by_parameter():
    """A parameter bound as `int` that the header declares `long long` (or a float bound as a pointer) would corrupt a call
    silently: every ctypes prototype of _lib.py is compared with the C declaration, type class by type class."""
    ctype_cat = {ctypes.c_int: "int", ctypes.c_longlong: "ll", ctypes.c_ulonglong: "ull", ctypes.c_float: "float",
                 ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr"}
    decls = _header_declarations()
    assert len(decls) >= 40
    checked = 0
    for name, cats in decls:
        proto = cm._lib.PROTOTYPES.get(name)
        if proto is None:
            continue
        bound = [ctype_cat.get(t, "ptr" if issubclass(t, ctypes._Pointer) else None) for t in proto]
        assert bound == cats, (name, cats, bound)
        checked += 1
    assert checked == len(cm._lib.PROTOTYPES)
    # struct cmrl_rollout_group: same field order and types as the ctypes Structure
    text = open(os.path.join(ROOT, "include", "cmrl_b200.h")).read()
    body = re.search(r"typedef struct cmrl_rollout_group\s*\{(.*?)\}\s*cmrl_rollout_group;", text, flags=re.S).group(1)
    body = re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    fields = []
    for stmt in body.split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        ptr = "*" in stmt
        base = "ptr" if ptr else ("ll" if stmt.startswith("long long") else "int")
        names = re.sub(r"^(const\s+)?(unsigned\s+)?(long long|int|float|void|uint8_t)\s*\**", "", stmt)
        for n_ in names.split(","):
            fields.append((n_.replace("*", "").strip(), base))
    got = [(n_, ctype_cat[t]) for n_, t in cm._lib.RolloutGroup._fields_]
    assert got == fields, (got, fields)


def test_header_is_plain_c_and_a_c_program_links_against_the_library():
    """The boundary is a C ABI: include/cmrl_b200.h compiles as C99 and as C++17, and a C program (no Python, no torch) links
    against libcmrl_b200.so, reads the ABI version and gets the documented error code / message for an invalid call."""
    import shutil
    import subprocess

    gcc, gxx = shutil.which("gcc"), shutil.which("g++")
    if gcc is None:
        pytest.skip("no C compiler")
    hdr = os.path.join(ROOT, "include", "cmrl_b200.h")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    if gxx is not None:
        subprocess.run([gxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr], check=True)
    libdir = os.path.dirname(cm._lib.LIB_PATH)
    src = os.path.join(ROOT, "examples", "abi_consumer.c")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "consumer")
        subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                        "-L", libdir, "-l:libcmrl_b200.so", "-Wl,-rpath," + libdir], check=True)
        res = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
        assert res.stdout.startswith("abi 1, launches ")


def test_no_cpu_fallback():
    m = cm.PlainTransition(5, 1, hid_size=8)
    with pytest.raises(cm._lib.CmrlB200Error):
        m.forward(torch.zeros(7, 3, 5), torch.zeros(7, 3, 1))
    with pytest.raises(cm._lib.CmrlB200Error):
        cm.EnsembleLinearLayer(3, 4, ensemble_num=2)(torch.zeros(2, 5, 3))


def test_oracle_is_test_infrastructure_only():
    """Nothing in the product package, and nothing on the measured (native) arm of bench.py, touches oracle/."""
    import glob
    import inspect
    import re

    for f in glob.glob(os.path.join(ROOT, "causal-mbrl_b200", "*.py")) + glob.glob(os.path.join(ROOT, "cmrl_b200", "*.py")):
        assert not re.search(r"^\s*(from|import)\s+oracle", open(f).read(), re.M), f
    sys.path.insert(0, ROOT)
    import bench

    for fn in (bench.run_native, bench.time_train, bench.dominant_kernel_roofline):
        assert not re.search(r"(from|import)\s+oracle", inspect.getsource(fn)), fn.__name__


def test_unsupported_activation_is_loud():
    with pytest.raises(NotImplementedError):
        cm.PlainTransition(5, 1, hid_size=8, activation_fn_cfg={"_target_": "torch.nn.Tanh"})


@pytest.mark.parametrize("cls,kw,fname", [
    ("PlainTransition", {}, "plain_transition.pth"),
    ("ExternalMaskTransition", {}, "external_mask_transition.pth"),
    ("PlainRewardMech", {}, "plain_reward_mech.pth"),
    ("PlainTerminationMech", {}, "plain_termination_mech.pth"),
])
def test_state_dict_layout_and_checkpoint_roundtrip(cls, kw, fname):
    m = getattr(cm, cls)(5, 1, hid_size=16, activation_fn_cfg=SILU, **kw)
    sd = m.state_dict()
    lead = (5, 7) if cls == "ExternalMaskTransition" else (7,)
    head = 10 if cls == "PlainTransition" else 2
    assert tuple(sd["hidden_layers.0.0.weight"].shape) == lead + (6, 16)
    assert tuple(sd["hidden_layers.3.0.bias"].shape) == lead + (1, 16)
    assert tuple(sd["mean_and_logvar.weight"].shape) == lead + (16, head)
    bounds = {"PlainTransition": (1, 5), "ExternalMaskTransition": (5, 1, 1, 1)}.get(cls, (1,))
    assert tuple(sd["min_logvar"].shape) == bounds and float(sd["min_logvar"].flatten()[0]) == -10.0
    assert float(sd["max_logvar"].flatten()[0]) == 0.5 and not m.min_logvar.requires_grad
    assert m.model_file_name == fname
    with tempfile.TemporaryDirectory() as d:
        m.save(d)
        assert os.listdir(d) == [fname]
        m2 = getattr(cm, cls)(5, 1, hid_size=16, activation_fn_cfg=SILU, **kw)
        m2.load(d)
        for k in sd:
            assert torch.equal(sd[k], m2.state_dict()[k])
        assert list(np.asarray(m2.elite_members)) == list(np.asarray(m.elite_members))
        if cls == "ExternalMaskTransition":
            assert torch.equal(m2.input_mask, m.input_mask) and "input_mask" in m.save_attr


@pytest.mark.skipif(not refshim.mounted(), reason="needs the reference tree (build container)")
@pytest.mark.parametrize("cls", ["PlainTransition", "ExternalMaskTransition", "PlainRewardMech", "PlainTerminationMech"])
def test_checkpoints_cross_load_with_the_reference(cls):
    """`.pth` files are part of the boundary (mlp.py:46-63): a checkpoint written by the unmodified reference loads into the
    drop-in class and one written here loads into the reference -- same file name, same keys, weights, elite set and (for the
    masked transition) input mask."""
    ref = refshim.load_reference()
    rng = np.random.default_rng(4)
    for writer, reader in ((ref, cm), (cm, ref)):
        torch.manual_seed(11)
        src = getattr(writer, cls)(5, 1, hid_size=16, activation_fn_cfg=SILU)
        dst = getattr(reader, cls)(5, 1, hid_size=16, activation_fn_cfg=SILU)
        src.set_elite_members([6, 1, 3, 0, 4])
        if cls == "ExternalMaskTransition":
            src.set_input_mask(torch.from_numpy((rng.uniform(size=(5, 6)) < 0.5).astype(np.float32)))
        assert src.model_file_name == dst.model_file_name
        with tempfile.TemporaryDirectory() as d:
            src.save(d)
            assert os.listdir(d) == [dst.model_file_name]
            blob = torch.load(os.path.join(d, dst.model_file_name), weights_only=False)
            assert sorted(blob) == sorted(dst.save_attr)
            dst.load(d)
        a, b = src.state_dict(), dst.state_dict()
        assert list(a) == list(b)
        for k in a:
            assert torch.equal(a[k], b[k]), k
        assert list(map(int, dst.elite_members)) == [6, 1, 3, 0, 4]
        if cls == "ExternalMaskTransition":
            assert torch.equal(dst.input_mask, src.input_mask)


@pytest.mark.skipif(not refshim.mounted(), reason="needs the reference tree (build container)")
def test_public_signatures_are_the_references():
    """The drop-in boundary is the Python class API: every public method (and `__init__`) of the reference's classes
    on the path exists here with the same parameter names, kinds, order and defaults; every property exists.  Additions are
    allowed only as trailing parameters WITH defaults (e.g. `subnet_range=None`)."""
    import inspect

    ref = refshim.load_reference()
    classes = ["PlainTransition", "ExternalMaskTransition", "ForwardEulerTransition", "PlainRewardMech", "PlainTerminationMech",
               "BaseDynamics", "PlainEnsembleDynamics", "ConstraintBasedDynamics", "VecFakeEnv", "TransitionIterator",
               "BootstrapIterator", "InteractionBatch", "EnsembleLinearLayer", "ParallelEnsembleLinearLayer", "EnsembleMLP",
               "TransitionConditionalMutualInformationTest"]

    def params(f):
        return [(n, p_.kind.name, None if p_.default is inspect.Parameter.empty else repr(p_.default))
                for n, p_ in inspect.signature(f).parameters.items()]

    checked = 0
    for name in classes:
        theirs, ours = getattr(ref, name), getattr(cm, name)
        for meth, fn in inspect.getmembers(theirs, predicate=inspect.isfunction):
            if meth.startswith("_") and meth != "__init__":
                continue
            mine = getattr(ours, meth, None)
            assert mine is not None, "%s.%s is missing" % (name, meth)
            a, b = params(fn), params(mine)
            assert b[:len(a)] == a, "%s.%s: %s != %s" % (name, meth, b, a)
            assert all(d is not None for _, _, d in b[len(a):]), "%s.%s: extra parameters need defaults" % (name, meth)
            checked += 1
        for prop, _ in inspect.getmembers(theirs, predicate=lambda x: isinstance(x, property)):
            assert hasattr(ours, prop), "%s.%s (property) is missing" % (name, prop)
    assert checked > 120
    # instance attributes, parameters and sub-modules the callers read (SURVEY 8b's attribute list)
    def public(o):
        names = set(n for n in vars(o) if not n.startswith("_"))
        if isinstance(o, torch.nn.Module):
            names |= set(o._parameters) | set(o._modules)
        return names

    for name in ("PlainTransition", "ExternalMaskTransition", "PlainRewardMech", "PlainTerminationMech",
                 "TransitionConditionalMutualInformationTest"):
        theirs, ours = (getattr(ns, name)(5, 1, hid_size=8, activation_fn_cfg=SILU) for ns in (ref, cm))
        assert public(theirs) <= public(ours), (name, public(theirs) - public(ours))
    t_ref, t_our = ref.PlainTransition(5, 1, hid_size=8), cm.PlainTransition(5, 1, hid_size=8)
    assert public(ref.ForwardEulerTransition(t_ref, 3)) <= public(cm.ForwardEulerTransition(t_our, 3))
    d_ref, d_our = ref.PlainEnsembleDynamics(t_ref, learned_reward=False), cm.PlainEnsembleDynamics(t_our, learned_reward=False)
    assert public(d_ref) <= public(d_our), public(d_ref) - public(d_our)
    e_ref, e_our = ref.VecFakeEnv(4, None, None), cm.VecFakeEnv(4, None, None)
    assert all(hasattr(e_our, n) for n in vars(e_ref)), [n for n in vars(e_ref) if not hasattr(e_our, n)]


def test_elite_bookkeeping_and_random_index():
    m = cm.PlainTransition(5, 1, hid_size=8)
    assert len(m.elite_members) == 5
    m.set_elite_members([0, 2, 3, 5, 6])
    m.set_elite_members(list(range(7)))  # ignored when all members are passed (mlp.py:31-34)
    assert m.elite_members == [0, 2, 3, 5, 6]
    a = m.get_random_index(100, np.random.default_rng(3))
    b = np.random.default_rng(3).choice([0, 2, 3, 5, 6], size=100)
    assert np.array_equal(a, b)


def test_interaction_batch():
    """tests/test_types.py:9-47 of the reference."""
    n = 12
    b = cm.InteractionBatch(np.zeros((n, 3)), np.zeros((n, 2)), np.zeros((n, 3)), np.zeros(n), np.zeros(n, bool))
    assert len(b) == n and len(b.as_tuple()) == 5
    sub = b[[1, 4, 5]]
    assert len(sub) == 3 and sub.batch_obs.shape == (3, 3)
    r = b.add_new_batch_dim(4)
    assert r.batch_obs.shape == (4, 3, 3) and r.batch_reward.shape == (4, 3) and r.batch_done.shape == (4, 3)
    with pytest.raises(ValueError):
        b.add_new_batch_dim(5)
    assert b.attrs == ["batch_obs", "batch_action", "batch_next_obs", "batch_reward", "batch_done"]


def _data(n, rng):
    return cm.InteractionBatch(rng.normal(size=(n, 3)).astype(np.float32), rng.normal(size=(n, 1)).astype(np.float32),
                               rng.normal(size=(n, 3)).astype(np.float32), rng.normal(size=n).astype(np.float32),
                               np.zeros(n, bool))


def test_iterators_shapes_and_coverage():
    rng = np.random.default_rng(0)
    data = _data(103, rng)
    it = cm.TransitionIterator(data, 32, shuffle_each_epoch=True, rng=np.random.default_rng(1))
    assert len(it) == 4
    seen = np.concatenate([b.batch_reward for b in it])
    assert np.array_equal(np.sort(seen), np.sort(data.batch_reward))  # a permutation, ragged last batch of 7
    bt = cm.BootstrapIterator(data, 32, 7, shuffle_each_epoch=True, permute_indices=False, rng=np.random.default_rng(2))
    batches = list(bt)
    assert [b.batch_obs.shape for b in batches] == [(7, 32, 3)] * 3 + [(7, 7, 3)]
    assert batches[0].batch_reward.shape == (7, 32) and batches[0].batch_done.dtype == bool
    mi, order = bt.device_plan()
    assert mi.shape == (7, 103) and np.array_equal(batches[1].batch_obs[3], data.batch_obs[mi[3, order[32:64]]])
    bp = cm.BootstrapIterator(data, 32, 7, permute_indices=True, rng=np.random.default_rng(2))
    assert all(np.array_equal(np.sort(row), np.arange(103)) for row in bp.member_indices)
    bp.toggle_bootstrap()
    assert next(iter(bp)).batch_obs.shape == (32, 3)


def test_prefetched_epoch_order_is_the_order_the_next_pass_would_have_drawn():
    """The device-resident epoch draws the next pass's permutation while the GPU runs the current one: the sequence of
    passes must not change, and a caller-supplied generator (other consumers may sit between two draws) is left alone."""
    data = _data(57, np.random.default_rng(0))
    a = cm.TransitionIterator(data, 16, shuffle_each_epoch=True)
    b = cm.TransitionIterator(data, 16, shuffle_each_epoch=True)
    a._rng, b._rng = np.random.default_rng(11), np.random.default_rng(11)
    orders_a, orders_b = [], []
    for k in range(4):
        iter(a)
        orders_a.append(a._order.copy())
        if k % 2 == 0:
            nxt = a.prefetch_order()
            assert nxt is not None and a.prefetch_order() is nxt  # drawn once
        iter(b)
        orders_b.append(b._order.copy())
    assert all(np.array_equal(x, y) for x, y in zip(orders_a, orders_b))
    # drawn on the helper thread: the next pass joins it and takes that draw; the callback saw the same array
    seen = []
    assert a.prefetch_order(background=True, then=seen.append) is None
    iter(a), iter(b)
    assert np.array_equal(a._order, b._order) and len(seen) == 1 and seen[0] is a._order
    shared = cm.TransitionIterator(data, 16, shuffle_each_epoch=True, rng=np.random.default_rng(3))
    assert shared.prefetch_order() is None
    assert cm.TransitionIterator(data, 16, shuffle_each_epoch=False).prefetch_order() is None


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted (GPU box)")
def test_iterators_match_reference_bit_for_bit():
    ref = refshim.load_reference()
    rng = np.random.default_rng(0)
    data = _data(257, rng)
    rdata = ref.InteractionBatch(*data.as_tuple())
    for permute in (False, True):
        a = cm.BootstrapIterator(data, 64, 7, shuffle_each_epoch=True, permute_indices=permute, rng=np.random.default_rng(5))
        b = ref.BootstrapIterator(rdata, 64, 7, shuffle_each_epoch=True, permute_indices=permute, rng=np.random.default_rng(5))
        for _ in range(2):  # two epochs: the shuffle consumes the generator identically
            for x, y in zip(a, b):
                for u, v in zip(x.as_tuple(), y.as_tuple()):
                    assert np.array_equal(u, v) and u.dtype == v.dtype and u.shape == v.shape


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("cls,kw", [("PlainTransition", {}), ("ExternalMaskTransition", {}), ("PlainRewardMech", {}),
                                     ("PlainTransition", {"activation_fn_cfg": None, "deterministic": True}),
                                     ("TransitionConditionalMutualInformationTest", {}),
                                     ("TransitionConditionalMutualInformationTest", {"activation_fn_cfg": None, "learn_logvar_bounds": True})])
def test_same_seed_gives_the_reference_initial_weights(cls, kw):
    """Constructing with the same torch / numpy seeds reproduces the reference's init (truncated normal
    resampling order, layers.py:8-28) and its random initial elites (mlp.py:27) exactly."""
    ref = refshim.load_reference()
    args = dict(hid_size=24, activation_fn_cfg=SILU)
    args.update(kw)
    torch.manual_seed(7)
    np.random.seed(7)
    a = getattr(cm, cls)(5, 1, **args)
    torch.manual_seed(7)
    np.random.seed(7)
    b = getattr(ref, cls)(5, 1, **args)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert list(a.elite_members) == list(b.elite_members)
    assert [p.requires_grad for p in a.parameters()] == [p.requires_grad for p in b.parameters()]
    assert a.model_file_name == b.model_file_name
    if hasattr(b, "input_mask"):
        assert torch.equal(a.input_mask.cpu(), b.input_mask.cpu())


def test_dynamics_host_logic_without_gpu():
    tr = cm.PlainTransition(3, 1, hid_size=8)
    rw = cm.PlainRewardMech(3, 1, hid_size=8)
    tm = cm.PlainTerminationMech(3, 1, hid_size=8)
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw, learned_termination=True, termination_mech=tm)
    assert dyn.learn_mech == ["transition", "reward_mech", "termination_mech"]
    assert dyn.get_variable_by_mech("reward_mech") == "batch_reward" and dyn.get_mach_by_variable("batch_next_obs") == "transition"
    assert dyn.total_epoch == {"transition": 0, "reward_mech": 0, "termination_mech": 0}
    rng = np.random.default_rng(0)
    n = 50
    buf = refshim.DuckReplayBuffer(rng.normal(size=(n, 3)).astype(np.float32), rng.normal(size=(n, 1)).astype(np.float32),
                                   rng.normal(size=(n, 3)).astype(np.float32), rng.normal(size=n).astype(np.float32), np.zeros(n, bool))
    tr_it, val_it = dyn.dataset_split(buf, validation_ratio=0.2, batch_size=16)
    assert tr_it.num_stored == 40 and val_it.num_stored == 10 and isinstance(tr_it, cm.BootstrapIterator)
    assert np.array_equal(val_it.transitions.batch_obs, buf.observations[40:, 0])  # tail = validation, no shuffle
    # elite selection and best-weight snapshot rules (plain_dynamics.py:113-142)
    best = torch.tensor([1.0, 2.0, 0.5, 4.0, 3.0, 0.1, 9.0])
    assert dyn.maybe_get_best_weights(best, best * 0.95, "transition", 0.1) is None
    snap = dyn.maybe_get_best_weights(best, torch.where(torch.arange(7) == 3, best * 0.5, best), "transition", 0.1)
    assert snap is not None and set(snap) == set(tr.state_dict())
    dyn.maybe_set_best_weights_and_elite(None, best, "transition")
    assert list(tr.elite_members) == [5, 2, 0, 1, 4]
    t = dyn.get_3d_tensor(np.ones((4, 3), np.float32), is_ensemble=False)
    assert tuple(t.shape) == (7, 4, 3) and t.stride(0) == 0  # broadcast view, never replicated
    assert tuple(dyn.get_3d_tensor(np.ones((7, 4), np.float32), is_ensemble=True).shape) == (7, 4, 1)


def test_constraint_based_dynamics_host_logic_matches_the_reference():
    """ConstraintBasedDynamics (constraint_based_dynamics.py:22-111): oracle / history mask attributes only for mechanisms that
    have an input mask, `set_oracle_mask`, and `learn` without an oracle mask failing in `set_input_mask(None)` exactly like the
    reference -- compared with the unmodified class where the reference tree is mounted."""
    def build(ns, extra):
        tr = ns.ExternalMaskTransition(4, 2, hid_size=8, **extra)
        rw = ns.PlainRewardMech(4, 2, hid_size=8, **extra)
        return ns.ConstraintBasedDynamics(tr, learned_reward=True, reward_mech=rw)

    rng = np.random.default_rng(0)
    n = 40
    buf = refshim.DuckReplayBuffer(rng.normal(size=(n, 4)).astype(np.float32), rng.normal(size=(n, 2)).astype(np.float32),
                                   rng.normal(size=(n, 4)).astype(np.float32), rng.normal(size=n).astype(np.float32), np.zeros(n, bool))
    mask = (rng.uniform(size=(4, 6)) < 0.5).astype(np.float32)
    variants = [(cm, {})]
    if refshim.mounted():
        variants.append((refshim.load_reference(), {}))
    seen = []
    for ns, extra in variants:
        dyn = build(ns, extra)
        attrs = sorted(a for a in vars(dyn) if a.endswith("_oracle_mask") or a.endswith("_history_mask"))
        assert attrs == ["transition_history_mask", "transition_oracle_mask"]
        assert dyn.transition_oracle_mask is None
        assert torch.equal(dyn.transition_history_mask, torch.ones(4, 6))
        with pytest.raises(AssertionError):
            dyn.set_oracle_mask("reward_mech", mask)  # the reward mech has no input mask
        with pytest.raises(ValueError):
            dyn.learn(buf, batch_size=16, longest_epoch=1, work_dir=None)  # to_tensor(None) (util.py:64-69)
        dyn.set_oracle_mask("transition", mask)
        assert torch.equal(dyn.transition_oracle_mask, torch.from_numpy(mask))
        assert torch.equal(dyn.transition.input_mask, torch.ones(4, 6))  # applied by learn(), not by set_oracle_mask
        seen.append((dyn.learn_mech, dict(dyn.total_epoch)))
    assert all(x == seen[0] for x in seen)


def test_fake_env_host_side_contract():
    env = cm.VecFakeEnv(8, None, None)
    assert env.num_envs == 8 and not env.has_set_up and env.reset() is None
    with pytest.raises(AssertionError):
        env.step_async(np.zeros((8, 1), np.float32))
        env.step_wait()
    full = np.arange(16, dtype=np.float32).reshape(8, 2)  # third argument: the pre-reset rows of the done envs, in done_idx order
    infos = env._make_infos(np.array([False, True] + [False] * 6), np.array([1, 3]), full[[1, 3]])
    assert [i["TimeLimit.truncated"] for i in infos] == [False, True] + [False] * 6
    assert "terminal_observation" in infos[1] and "terminal_observation" in infos[3] and "terminal_observation" not in infos[0]
    env.infos_mode = "shared"
    infos2 = env._make_infos(np.array([False, True] + [False] * 6), np.array([1, 3]), full[[1, 3]])
    assert [bool(i["TimeLimit.truncated"]) for i in infos2] == [False, True] + [False] * 6
    assert np.array_equal(infos2[3]["terminal_observation"], [6.0, 7.0]) and infos2[0] is infos2[2]
    # shared mode: a step without done envs hands out one cached list; a step with done envs never writes into it
    none_done = env._make_infos(np.zeros(8, bool), np.array([], dtype=np.int64), None)
    assert none_done is env._make_infos(np.zeros(8, bool), np.array([], dtype=np.int64), None)
    assert len(none_done) == 8 and all(not i["TimeLimit.truncated"] and "terminal_observation" not in i for i in none_done)
    assert infos2 is not none_done and "terminal_observation" not in none_done[1]


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted (GPU box)")
def test_batched_reset_draws_equal_reference_per_env_draws():
    """One get_init_obs_fn(n) call == n calls of size 1 for generator-backed callbacks (DESIGN.md decision 3)."""
    a, b = np.random.default_rng(3), np.random.default_rng(3)
    x = a.normal(size=(5, 4)).astype(np.float32)
    y = np.concatenate([b.normal(size=(1, 4)).astype(np.float32) for _ in range(5)])
    assert np.array_equal(x, y)


def test_draw_ahead_buffers_are_trusted_only_after_a_replay_of_the_same_graph_family():
    """VecFakeEnv._ahead_before_replay: the transition draws made one step ahead are reused only if the previous
    counter-advancing event was a replay of the same graph family; anything else draws them before the replay."""
    env = cm.VecFakeEnv(4, None, None)
    drawn = []
    env._ahead_draw = lambda stream_id, mechs=(): drawn.append(stream_id)  # the kernels are not under test here
    predict = {"t_id": 5, "n": 3, "family": ("predict", "key", 5, 3)}
    step = {"t_id": 5, "n": 4, "family": ("step", 1234, 5, 4)}
    env._ahead_before_replay(predict)  # first replay: nothing prepared yet
    env._ahead_before_replay(predict)  # steady state
    env._ahead_before_replay(dict(predict))  # the second graph of the family (other pinned destination buffer)
    assert drawn == [5]
    env._ahead_before_replay(step)  # another step flavour advanced the counter by its own n
    env._ahead_before_replay(predict)
    assert drawn == [5, 5, 5]
    env._ahead_before_replay(None)  # a graph without draw-ahead
    env._ahead_before_replay(predict)
    assert drawn == [5, 5, 5, 5]
    env._ahead_family = None  # what an eager transition prepare / seed() / invalidate_step_graph() leave behind
    env._ahead_before_replay(predict)
    assert drawn == [5] * 5


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("n,ratio,full", [(50, 0.2, True), (37, 0.33, True), (64, 0.0, True), (257, 0.2, False)])
def test_dataset_split_and_elite_selection_match_the_reference(n, ratio, full):
    """base_dynamics.py:123-156 (train / validation split of an SB3 buffer) and plain_dynamics.py:113-142 (best-weights
    snapshot rule, elite set) against the live reference on the same inputs."""
    ref = refshim.load_reference()
    rng = np.random.default_rng(n)
    buf = refshim.DuckReplayBuffer(rng.normal(size=(n, 3)).astype(np.float32), rng.normal(size=(n, 1)).astype(np.float32),
                                   rng.normal(size=(n, 3)).astype(np.float32), rng.normal(size=n).astype(np.float32), rng.uniform(size=n) < 0.1)
    if not full:  # a partly filled ring buffer: only the first `pos` rows count
        buf.full, buf.pos = False, n - 40
    torch.manual_seed(0)
    mine = cm.PlainEnsembleDynamics(cm.PlainTransition(3, 1, hid_size=8), learned_reward=False)
    theirs = ref.PlainEnsembleDynamics(ref.PlainTransition(3, 1, hid_size=8), learned_reward=False)
    (a_tr, a_val), (b_tr, b_val) = mine.dataset_split(buf, ratio, 16), theirs.dataset_split(buf, ratio, 16)
    assert (a_val is None) == (b_val is None) and a_tr.num_stored == b_tr.num_stored and len(a_tr) == len(b_tr)
    for x, y in zip(a_tr.transitions.as_tuple(), b_tr.transitions.as_tuple()):
        assert np.array_equal(x, y) and x.dtype == y.dtype
    if a_val is not None:
        assert len(a_val) == len(b_val)
        for x, y in zip(a_val.transitions.as_tuple(), b_val.transitions.as_tuple()):
            assert np.array_equal(x, y) and x.dtype == y.dtype
    for _ in range(6):
        best = torch.from_numpy(rng.uniform(0.1, 2.0, size=7).astype(np.float32))
        val = best * torch.from_numpy(rng.uniform(0.97, 1.02, size=7).astype(np.float32))
        assert (mine.maybe_get_best_weights(best, val, "transition", 0.01) is None) == \
            (theirs.maybe_get_best_weights(best, val, "transition", 0.01) is None)
        mine.maybe_set_best_weights_and_elite(None, val, "transition")
        theirs.maybe_set_best_weights_and_elite(None, val, "transition")
        assert [int(e) for e in mine.transition.elite_members] == [int(e) for e in theirs.transition.elite_members]


def test_weights_token_sees_torch_writes_and_fused_optimiser_steps(monkeypatch):
    """Caches derived from the weights (packed fp16 copies, captured step graphs) are keyed by EnsembleMLP.weights_token():
    it must move on in-place torch writes (version counters) AND on the fused Adam kernel, which writes through raw
    pointers -- a rollout after `learn()` must never see the weights from before it."""
    torch.manual_seed(0)
    tr = cm.PlainTransition(3, 1, hid_size=8)
    rw = cm.PlainRewardMech(3, 1, hid_size=8)
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw)
    t0 = tr.weights_token()
    assert tr.weights_token() == t0  # stable while nothing changes
    tr.load_state_dict({k: v.clone() for k, v in tr.state_dict().items()})
    t1 = tr.weights_token()
    assert t1 > t0
    with torch.no_grad():
        next(tr.parameters()).mul_(1.0)
    t2 = tr.weights_token()
    assert t2 > t1
    # the fused optimiser: kernels replaced by no-ops (no GPU here), the bookkeeping is what is under test
    optim = dyn.transition_optimizer
    flat = torch.zeros(4)
    monkeypatch.setattr(optim, "_arena", lambda: (flat, flat))
    monkeypatch.setattr(cm.ops, "adam_l2_step", lambda *a, **k: None)
    monkeypatch.setattr(cm.ops, "adam_l2_step_dev", lambda *a, **k: None)
    optim._exp_avg = optim._exp_avg_sq = flat
    optim.step()
    t3 = tr.weights_token()
    assert t3 > t2
    monkeypatch.setattr(optim, "prepare_graphable", lambda: None)
    optim._step_dev = flat
    optim.step_graphable()
    assert tr.weights_token() > t3 and rw.weights_token() == rw.weights_token()
    # the key the fake env compares before it replays a captured step graph
    env = cm.VecFakeEnv(4, None, None)
    env.dynamics, env.precision, env.deterministic, env.penalty_coeff, env.max_episode_steps = dyn, "f16", False, 0.0, 10
    k0 = env._graph_state_key(False)
    assert env._graph_state_key(False) == k0 and env._step_graph_state()[:len(k0)] == k0
    optim._mark_weights_changed()  # what a replayed training graph reports per step
    k1 = env._graph_state_key(False)
    assert k1 != k0 and k1[4] == k0[4]  # the transition's entry moved, the reward mech's did not
    tr.set_elite_members([0, 1, 2, 3, 4])
    assert env._graph_state_key(False) != k1
    assert env._graph_state_key(True) != env._graph_state_key(False)


def test_round2_host_rules():
    """Small host-side rules added in round 2: net shard ranges cover the sub-networks exactly once in contiguous near-equal
    shares; the per-layer training GEMMs pick the tensor-core tier by tile count as well as by rows; mechanisms are grouped
    into one rollout launch only where that was measured to pay."""
    for P in (1, 5, 100):
        for world in (1, 2, 3, 8):
            if world > P:
                continue
            ranges = [cm.net_shard_range(P, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == P
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
    tier = cm.ops._train_tier
    assert tier("fwd", 256, 7, 200, 200) == 0 and tier("fwd", 4096, 7, 200, 200) == 3          # few nets: by rows
    assert tier("fwd", 256, 1600, 512, 512) == 3 and tier("dx", 256, 1600, 512, 512) == 3        # many nets: by tiles
    assert tier("fwd", 256, 1600, 512, 2) == 0 and tier("fwd", 256, 1600, 120, 512) == 0         # narrow layers stay on FFMA
    assert tier("dw", 256, 7, 200, 200) == 0 and tier("dw", 2048, 7, 200, 200) == 3 and tier("dw", 256, 1600, 512, 512) == 3
    keep = cm.ops.TRAIN_GEMM_TIER
    try:
        cm.ops.TRAIN_GEMM_TIER = 0
        assert tier("fwd", 1 << 20, 1600, 512, 512) == 0  # an explicit tier overrides the rule
    finally:
        cm.ops.TRAIN_GEMM_TIER = keep
    env = cm.VecFakeEnv(100_000, None, None)
    env.learned_reward, env.learned_termination = True, False
    assert env.group_mechs == "auto" and not env._group_on()      # two mechanisms, many envs: one launch per mechanism
    env.learned_termination = True
    assert env._group_on()                                          # three mechanisms
    small = cm.VecFakeEnv(1000, None, None)
    small.learned_reward, small.learned_termination = True, False
    assert small._group_on()                                        # few envs: launch bound
    small.group_mechs = False
    assert not small._group_on()


def test_external_mask_shard_is_the_full_models_slice_on_the_host():
    """A net shard built with subnet_range has the full model's interface restricted to its dimensions (no kernels involved)."""
    sh = cm.ExternalMaskTransition(6, 2, hid_size=8, subnet_range=(2, 5))
    assert sh.is_net_shard and sh.subnet_range == (2, 5) and sh._head_dim == 3 and sh._parallel_num() == 3 and sh._grad_dims() == 6
    assert [tuple(p.shape) for p in sh.parameters()][-2:] == [(3, 7, 8, 2), (3, 7, 1, 2)]  # head weight / bias of three sub-networks
    mask = np.arange(6 * 8, dtype=np.float32).reshape(6, 8)
    sh.set_input_mask(torch.from_numpy(mask))
    assert np.array_equal(sh._kernel_mask().numpy(), mask[2:5]) and np.array_equal(sh.input_mask.numpy(), mask)
    tgt = torch.arange(24.0).view(2, 2, 6)
    assert torch.equal(sh.target_columns(tgt), tgt[..., 2:5])
    full = cm.ExternalMaskTransition(6, 2, hid_size=8)
    assert not full.is_net_shard and full.target_columns(tgt) is tgt and full._nll_reg() == pytest.approx(0.01 * 6 * 10.5)
    assert sh._nll_reg() == pytest.approx(full._nll_reg())  # the loss constant of the FULL model
    with pytest.raises(AssertionError):
        cm.ExternalMaskTransition(6, 2, hid_size=8, subnet_range=(2, 5), learn_logvar_bounds=True)


def test_peer_exchange_partition_covers_every_gradient_once():
    """csrc/dp_sync.cu cuts an arena of n floats into W slices of S floats and every slice into 148 CTA sub-ranges of T floats
    (both multiples of 4, the last ones clipped to n).  Restated here: every float belongs to exactly one (slice, CTA), every
    16-byte vector lies inside one sub-range, and the buffers sized for the communicator's capacity hold any smaller arena."""
    ctas = 148
    for world in (1, 2, 3, 8, 16):
        for n in (1, 3, 4, 5, 1000, 868072, 868075, 12_345_679):
            S = ((n + world - 1) // world + 3) // 4 * 4
            T = ((S + ctas - 1) // ctas + 3) // 4 * 4
            assert S % 4 == 0 and T % 4 == 0 and S * world >= n and T * ctas >= S
            covered = 0
            prev_end = 0
            for q in range(world):
                for b in range(ctas):
                    lo, hi = b * T, min((b + 1) * T, S)
                    if hi <= lo:
                        continue
                    g_lo, g_hi = min(n, q * S + lo), min(n, q * S + hi)
                    assert g_lo == prev_end or g_lo == n  # contiguous, in order, no overlap
                    covered += g_hi - g_lo
                    prev_end = max(prev_end, g_hi)
                    assert lo % 4 == 0  # vectors never straddle two CTAs
            assert covered == n and prev_end == n
            cap = n + 12345
            slice_cap = ((cap + world - 1) // world + 3) // 4 * 4
            assert S <= slice_cap
