"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, configuration handling mirrors the reference constructor, the product path refuses to run
without a GPU (no CPU fallback), and the small exactness lemmas the kernels rely on hold."""
import ctypes as C
import os
import re
from fractions import Fraction

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _state(**over):
    st = dict(type=2, add_action=True, add_reward=False, add_index=False, add_velocity=False,
              action_index="binary", piggybacking=False, add_position=False, add_positional_dist=False,
              add_positional_dist_piggy=True, add_positional_dist_type=2, add_channel_obs=False, num_bins=20)
    st.update(over)
    return st


def test_library_exports_every_declared_symbol():
    from diral_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "diral_env.h")).read()
    declared = set(re.findall(r"\b(diral_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found in include/diral_env.h"
    assert declared == set(_lib.SYMBOLS), "binding and header disagree: %s" % (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), "libdiral_env.so does not export %s" % name
    assert lib.diral_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_the_header():
    """ctypes mirrors must have the C layout: sizes are checked against a tiny C program."""
    import subprocess
    import tempfile
    from diral_b200._lib import DiralBuffers, DiralCfg
    src = '#include <stdio.h>\n#include "diral_env.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(diral_cfg), ' \
          'sizeof(diral_buffers), __builtin_offsetof(diral_cfg, sentinel), __builtin_offsetof(diral_buffers, trace_len));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).split()
    assert [int(x) for x in out] == [C.sizeof(DiralCfg), C.sizeof(DiralBuffers), DiralCfg.sentinel.offset,
                                     DiralBuffers.trace_len.offset]


def test_state_space_matches_reference_formula_and_oracle():
    from diral_b200 import _lib, cfg_from_kwargs
    from oracle.c_oracle import cfg_from_kwargs as orc_cfg, lib as orc_lib
    lib = _lib.load()
    rs = np.random.RandomState(3)
    for _ in range(200):
        st = _state(add_action=bool(rs.randint(2)), add_reward=bool(rs.randint(2)), add_index=bool(rs.randint(2)),
                    add_velocity=bool(rs.randint(2)), action_index=["binary", "real"][rs.randint(2)],
                    add_position=bool(rs.randint(2)), add_positional_dist=bool(rs.randint(2)),
                    add_positional_dist_piggy=bool(rs.randint(2)), add_channel_obs=bool(rs.randint(2)),
                    num_bins=int(rs.randint(1, 64)))
        kw = dict(num_users=int(rs.randint(2, 64)), num_channels=int(rs.randint(1, 40)),
                  enable_fingerprint=bool(rs.randint(2)), State=st)
        cfg = cfg_from_kwargs(4, 0, kw)
        s = lib.diral_state_space(C.byref(cfg))
        n, r, b = kw["num_users"], kw["num_channels"], st["num_bins"]
        want = ((r if st["action_index"] == "binary" else 1) if st["add_action"] else 0) + (r if st["add_channel_obs"] else 0) \
            + st["add_reward"] + st["add_index"] + st["add_velocity"] + 2 * st["add_position"] \
            + (n - 1) * st["add_positional_dist"] + 2 * kw["enable_fingerprint"] + b * st["add_positional_dist_piggy"]
        assert s == want                                        # reference test_env.py:49-85
        assert s == orc_lib().orc_state_space(C.byref(orc_cfg(**kw)))
    # shipped toy config: S = R + B = 23
    cfg = cfg_from_kwargs(1, 0, dict(num_users=4, num_channels=3, State=_state()))
    assert lib.diral_state_space(C.byref(cfg)) == 23


def test_constructor_defaults_and_rejections():
    from diral_b200 import cfg_from_kwargs
    cfg = cfg_from_kwargs(8, 16, dict(State=_state()))
    # TestEnv.__init__ defaults (reference test_env.py:12-24)
    assert (cfg.N, cfg.R, cfg.L, cfg.C, cfg.W, cfg.reward_design) == (3, 3, 200.0, 1.0, 500.0, 1)
    assert (cfg.E, cfg.env0, cfg.age_threshold, cfg.sentinel) == (8, 16, 20, 100000.0)
    assert not cfg.mobility and not cfg.toy and not cfg.fingerprint
    with pytest.raises(ValueError):
        cfg_from_kwargs(1, 0, dict())                           # State block is mandatory in the reference
    with pytest.raises(ValueError):
        cfg_from_kwargs(1, 0, dict(State=_state(piggybacking=True)))
    with pytest.raises(ValueError):
        cfg_from_kwargs(1, 0, dict(proportional_fair=True, State=_state()))
    with pytest.raises(ValueError):
        cfg_from_kwargs(1, 0, dict(State=_state(action_index="gray")))
    bad = _state(); del bad["num_bins"]
    with pytest.raises(ValueError):
        cfg_from_kwargs(1, 0, dict(State=bad))


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from diral_b200 import TestEnv, _lib, cfg_from_kwargs
    with pytest.raises(RuntimeError, match="CUDA"):
        TestEnv(num_envs=2, device="cuda", num_users=4, num_channels=3, State=_state())
    with pytest.raises(RuntimeError, match="CUDA"):
        TestEnv(num_envs=2, device="cpu", num_users=4, num_channels=3, State=_state())
    lib = _lib.load()
    cfg = cfg_from_kwargs(2, 0, dict(num_users=4, num_channels=3, State=_state()))
    h = C.c_void_p()
    rc = lib.diral_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and not h.value and b"cuda" in lib.diral_last_error().lower()
    # the RealNeS-side helpers and the replay ring are device-only as well
    import numpy as np
    from diral_b200 import realness
    from diral_b200.replay import Memory
    z = np.zeros((2, 3), dtype=np.float32)
    with pytest.raises(RuntimeError, match="CUDA"):
        realness.pack_tables(z, z, z.astype(np.int32), z.astype(np.int32))
    with pytest.raises(RuntimeError, match="CUDA"):
        realness.SemiPersistentScheduling(4, 20, -97.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        Memory(8, agents=4, state_space=3)


def test_wire_entry_layout_matches_the_c_struct():
    """realness.WIRE_DTYPE is diral_wire_entry (include/diral_env.h): 16 bytes, MA_NeighborTableEntry field order."""
    from diral_b200 import realness
    assert realness.WIRE_DTYPE.itemsize == 16
    assert realness.WIRE_DTYPE.names == ("pos_x", "pos_y", "seq_num", "last_update")
    assert [realness.WIRE_DTYPE.fields[n][1] for n in realness.WIRE_DTYPE.names] == [0, 4, 8, 12]
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "diral_env.h")).read()
    assert "float   pos_x, pos_y;" in header and "int32_t seq_num, last_update;" in header


def test_product_code_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "diral_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/diral_oracle.c", "").replace("oracle/", "") or f == "_never_", \
                    "%s mentions the oracle" % f


def test_size_helpers():
    from diral_b200 import _lib, cfg_from_kwargs
    lib = _lib.load()
    cfg = cfg_from_kwargs(4096, 0, dict(num_users=32, num_channels=20, State=_state()))
    n = lib.diral_state_bytes(C.byref(cfg))
    tables = 4096 * 32 * 32 * (4 + 4 + 8 + 4)
    assert tables < n < tables * 1.6
    assert lib.diral_scratch_bytes(C.byref(cfg)) == 0
    big = cfg_from_kwargs(16, 0, dict(num_users=256, num_channels=128, State=_state()))
    assert lib.diral_scratch_bytes(C.byref(big)) == 16 * 256 * 260 * 4      # keys no longer fit shared memory (row stride 260 words)


def test_count_over_len_division_is_exact():
    """The kernels form VPD = count / len as  q0 = c*r; q = fma(fma(-q0, m, c), r, q0)  with
    r = RN(1/m): it must equal RN(c/m) (what float32(np.float64(c)/m) gives).  Sampled here, checked
    exhaustively for all 0 <= c <= m < 1024 when the kernel was written."""
    f32 = np.float32

    def rn32(fr):
        x = f32(float(fr))
        cands = [x, np.nextafter(x, f32(np.inf)), np.nextafter(x, f32(-np.inf))]
        return min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), int(np.asarray(c).view(np.uint32)) & 1))

    rs = np.random.RandomState(0)
    ms = sorted(set([1, 2, 3, 5, 7, 19, 31, 127, 255, 1023] + [int(v) for v in rs.randint(1, 1024, 30)]))
    for m in ms:
        r = rn32(Fraction(1, m))
        for c in sorted(set([0, 1, m // 2, m - 1, m] + [int(v) for v in rs.randint(0, m + 1, 8)])):
            q0 = rn32(Fraction(c) * Fraction(float(r)))
            rem = rn32(Fraction(c) - Fraction(float(q0)) * m)
            q = rn32(Fraction(float(rem)) * Fraction(float(r)) + Fraction(float(q0)))
            assert q == f32(np.float64(c) / np.float64(m)), (c, m)


def test_signed_distance_identity():
    """On a flat highway the kernels use s = xpos - own_x for  sign * sqrt((own_x - xpos)**2):
    identical in float64, including the reference's libm pow (network.py:549-555)."""
    import math
    rs = np.random.RandomState(1)
    for _ in range(20000):
        x1, x2 = float(rs.uniform(0, 8000)), float(rs.uniform(0, 8000))
        d = math.sqrt((x2 - x1) ** 2 + (0.0 - 0.0) ** 2)
        s_ref = d * (1 if x1 - x2 > 0.0 else -1)
        assert s_ref == x1 - x2 or (s_ref == 0 and x1 - x2 == 0)


# the EnvironmentTest blocks of the six shipped YAML files (configs/4ue_3r_toy/*.yaml): they differ in num_bins only
_SHIPPED_ENVIRONMENT_TEST = dict(
    congestion_test=True, load_positions=False,
    load_file_pos="save_results/realness/drqn/config_realness_pos_store/positions_2000.npy", num_channels=3, num_users=4,
    mobility=True, mobility_vary=False, highway_length=100, enable_fingerprint=False, reward_design=2,
    communication_range=250,
    State=dict(type=2, add_action=True, add_reward=False, add_index=False, add_velocity=False, action_index="binary",
               piggybacking=False, add_position=False, add_positional_dist=False, add_positional_dist_piggy=True,
               add_positional_dist_type=2, add_channel_obs=False, num_bins=20))


@pytest.mark.parametrize("bins", [10, 20, 20, 20, 20, 40])     # b10_dis_07, b20_dis_03/05/07/95, b40_dis_07
def test_every_shipped_yaml_block_is_accepted(bins):
    """cfg_from_kwargs takes the shipped EnvironmentTest blocks verbatim (TestEnv(**cfg["EnvironmentTest"]),
    main_test.py:46) and yields the state space the reference computes (test_env.py:49-85: R + num_bins)."""
    from diral_b200 import _lib, cfg_from_kwargs
    kw = dict(_SHIPPED_ENVIRONMENT_TEST, State=dict(_SHIPPED_ENVIRONMENT_TEST["State"], num_bins=bins))
    cfg = cfg_from_kwargs(8, 0, kw)
    assert (cfg.N, cfg.R, cfg.B, cfg.toy, cfg.mobility, cfg.L, cfg.C, cfg.W) == (4, 3, bins, 1, 1, 100.0, 250.0, 500.0)
    assert _lib.load().diral_state_space(C.byref(cfg)) == 3 + bins


@pytest.mark.parametrize("state,extra,threads", [
    (dict(), {}, 1), (dict(), {}, 4), (dict(), {}, -3),
    (dict(add_channel_obs=True, add_reward=True, add_index=True, add_position=True, add_velocity=True),
     dict(enable_fingerprint=True), 3),
    (dict(action_index="real", num_bins=7, add_reward=True), dict(num_channels=5, num_users=13), 2),
    (dict(add_action=False, num_bins=12), dict(num_channels=9), -2),
    (dict(add_positional_dist_piggy=False, add_channel_obs=True), dict(num_channels=8), 2),
    # layouts of the four-agents-per-step AVX-512 path: record block in two vectors, rows shorter than one vector,
    # the widest block it takes, one-hot only
    (dict(num_bins=28), dict(num_channels=12), 1), (dict(num_bins=8), dict(num_channels=4), 3),
    (dict(num_bins=32), dict(num_channels=32), 2), (dict(add_positional_dist_piggy=False), dict(num_channels=16), 1),
])
def test_host_row_assembly_matches_obtain_state_layout(state, extra, threads):
    """diral_expand_state_host -- the host half of the compact host format -- against the row layout of
    TestEnv.obtain_state (test_env.py:527-583) built with NumPy in float64 and rounded to float32 once."""
    from diral_b200 import _lib, cfg_from_kwargs
    lib = _lib.load()
    kw = dict(dict(num_users=32, num_channels=20, highway_length=800, reward_design=2, communication_range=250,
                   mobility=True, bin_range=500), **extra)
    st = _state(**state)
    kw["State"] = st
    E = 97
    cfg = cfg_from_kwargs(E, 0, kw)
    N, R, B = cfg.N, cfg.R, cfg.B
    A, S = E * N, lib.diral_state_space(C.byref(cfg))
    rs = np.random.RandomState(5)
    act = rs.randint(0, R, A).astype(np.int32); act[3] = R + 2; act[4] = -1          # clamped like the kernels do
    counts = (rs.randint(0, 4, (A, B)) * (rs.rand(A, 1) < 0.9)).astype(np.uint8)
    rews = rs.randn(A).astype(np.float32); obs = (rs.rand(A, R) * 300).astype(np.float32)
    px = rs.rand(A) * 800; py = rs.randint(0, 3, A).astype(np.float64); vel = rs.rand(A) + 1.1
    raw = np.full(A * S + 16, np.nan, np.float32)              # 64-byte aligned rows (the widest vector path needs them)
    off = (-raw.ctypes.data % 64) // 4
    out = raw[off:off + A * S].reshape(A, S)
    rc = lib.diral_expand_state_host(C.byref(cfg), A, act.ctypes.data, counts.ctypes.data, rews.ctypes.data, obs.ctypes.data,
                                     px.ctypes.data, py.ctypes.data, vel.ctypes.data, 3.0, 0.25, threads, out.ctypes.data)
    assert rc == 0, lib.diral_last_error()
    cols = []
    a = np.clip(act, 0, R - 1)
    if st["add_action"]:
        cols.append(np.eye(R)[a] if st["action_index"] == "binary" else a[:, None].astype(np.float64))
    if st["add_channel_obs"]:
        cols.append(obs.astype(np.float64))
    if st["add_positional_dist_piggy"]:
        m = counts.sum(1, keepdims=True).astype(np.float64)
        cols.append(np.where(m > 0, counts / np.maximum(m, 1), 0.0))
    if st["add_reward"]:
        cols.append(rews[:, None].astype(np.float64))
    if st["add_index"]:
        cols.append((np.arange(A) % N + 1)[:, None].astype(np.float64))
    if st["add_position"]:
        cols += [(px / cfg.L)[:, None], (py / 2)[:, None]]
    if st["add_velocity"]:
        cols.append(vel[:, None])
    if kw.get("enable_fingerprint"):
        cols += [np.full((A, 1), 3.0), np.full((A, 1), 0.25)]
    ref = np.concatenate(cols, axis=1).astype(np.float32)
    assert ref.shape == out.shape and (out == ref).all()


def test_host_pool_queue_flags_and_abort():
    """The row-assembly pool behind diral_step_host / diral_step_host_begin (diral_host.cpp) as a plain C++ program:
    five jobs queued at once and released by 'device' flags in a scrambled order, a publish()-released job, an
    aborted job whose flags never come, more jobs than ring slots -- rows and rewards equal single-threaded
    expand_rows throughout (tests/cpp/host_pool_test.cpp)."""
    import subprocess
    import tempfile
    csrc = os.path.join(ROOT, "diral_b200", "csrc")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "host_pool_test")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", csrc, os.path.join(ROOT, "tests", "cpp", "host_pool_test.cpp"),
                               os.path.join(csrc, "diral_host.cpp"), "-lpthread", "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "host pool ok" in out.stdout, out.stdout + out.stderr


def test_host_row_assembly_rejects_unfused_state_blocks():
    from diral_b200 import _lib, cfg_from_kwargs
    lib = _lib.load()
    kw = dict(num_users=8, num_channels=3, highway_length=200, State=_state(add_positional_dist_type=1))
    cfg = cfg_from_kwargs(2, 0, kw)
    out = np.zeros((16, 23), np.float32); act = np.zeros(16, np.int32); cnt = np.zeros((16, 20), np.uint8)
    rc = lib.diral_expand_state_host(C.byref(cfg), 16, act.ctypes.data, cnt.ctypes.data, None, None, None, None, None,
                                     0.0, 1.0, 1, out.ctypes.data)
    assert rc == -5 and b"compact" in lib.diral_last_error()


def test_reference_install_recipe_and_timing_harness():
    """oracle/install_ref.py places the unmodified env files under oracle/_ref/ (git-ignored); oracle/ref_python.py
    imports them with the two shims and steps the real TestEnv.  Skipped where neither the reference nor an
    installed copy exists."""
    from oracle import install_ref, ref_python
    if not install_ref.install() and not ref_python.available():
        pytest.skip("no /root/reference and no oracle/_ref on this machine")
    res = ref_python.time_reference(dict(_SHIPPED_ENVIRONMENT_TEST), seconds=0.2, warm=2, processes=2)
    assert res["one_core"] > 0 and res["all_cores_P_processes"] > 0 and res["P"] == 2
    ignored = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in ignored


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's keys,
    the C port as `value`, the unmodified Python reference beside it when oracle/_ref is installed."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "agent-steps/s" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("configs[2]") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["cpu_model"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    rp = cb["reference_python"]
    assert "unavailable" in rp or (rp["one_core"] > 0 and rp["all_cores_P_processes"] > 0 and rp["P"] >= 1)
