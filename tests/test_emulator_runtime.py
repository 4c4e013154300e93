"""CPU: the host warp emulator itself (tests/emu) on kernels with closed-form results -- shuffles,
votes, match, block barriers with exited threads, static / dynamic shared memory, atomics across blocks,
two-phase lane-pair updates, partial-mask __syncwarp, named barriers, stream capture and graph replay.
If the emulator mis-modelled one of these, the emulated parity tests would prove nothing."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


@pytest.fixture(scope="module")
def selftest():
    so = os.path.join(HERE, "build", "selftest.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    cmd = ["g++", "-O1", "-g1", "-std=c++17", "-fPIC", "-shared", "-fno-extern-tls-init", "-Wno-unknown-pragmas", f"-I{HERE}",
           os.path.join(HERE, "selftest.cpp"), os.path.join(HERE, "emu_runtime.cpp"), "-o", so, "-lpthread"]
    try:
        subprocess.run(cmd, check=True, capture_output=True)
    except (OSError, subprocess.CalledProcessError) as e:
        pytest.skip(f"cannot build the emulator self-test: {getattr(e, 'stderr', e)}")
    lib = C.CDLL(so)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(5000)
    sh = np.zeros(2 * 64 * 8)
    total = C.c_double(0.0)
    sync = np.zeros(64 + 16)
    counters = np.array([-1, 1, 0], dtype=np.int32)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    rc = lib.gwi_emu_selftest(dp(sh), dp(x), x.size, C.byref(total), dp(sync), counters.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return dict(sh=sh.reshape(2, 64, 8), x=x, total=total.value, sync=sync, counters=counters)


def test_shuffles_votes_match(selftest):
    sh = selftest["sh"]
    for b in range(2):
        for t in range(64):
            lane, warp = t % 32, t // 32
            o = sh[b, t]
            first = 32 * warp + 1
            assert o[0] == sum(range(first, first + 32))  # butterfly sum of the warp
            assert o[1] == (t + 1 - 3 if lane >= 3 else t + 1)  # shfl_up keeps the own value at the edge
            assert o[2] == (t + 1 + 5 if lane + 5 <= 31 else t + 1)
            assert o[3] == 32 * warp + 7 + 1
            assert int(o[4]) == sum(1 << l for l in range(32) if l % 3 == 0)
            assert o[5] == 3.0  # any = 1, all = 1
            assert int(o[6]) == 0xFF << (8 * (lane // 8))
            assert o[7] == warp


def test_block_reduction_with_exited_threads_and_cross_block_atomics(selftest):
    assert abs(selftest["total"] - selftest["x"].sum()) < 1e-9


def test_lane_pair_phases_partial_syncwarp_named_barriers_and_graph_replay(selftest):
    sync, counters = selftest["sync"], selftest["counters"]
    # slot[l] = (l+1) + (l+17) after the two phases; lanes 0..7 add 100 each
    expect = sum((l + 1) + (l + 17) + 100.0 for l in range(8))
    assert np.all(sync[:32] == expect)
    for blk in range(4):
        for pair in range(4):
            assert sync[64 + blk * 4 + pair] == 1000.0 + 2 * pair
    assert counters[0] == 31 and counters[1] == -31
    assert counters[2] == 2 * 32  # the captured launch was replayed twice (4 blocks x 8 warps each)
