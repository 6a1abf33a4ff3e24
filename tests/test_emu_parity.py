"""The `gpu` parity tests, re-run WITHOUT a GPU against a lock-step CPU emulation build of the product's
own CUDA sources (tests/cuemu: every CUDA thread is a fiber, barriers and warp primitives block until
the participants arrive, shared memory is poisoned per CTA, device allocations carry canaries).

Test infrastructure only.  It checks kernel LOGIC in the authoring container (indexing, barrier
placement, warp exchanges, launch order, graph replay, shared-memory opt-in limits); the real `-m gpu`
run on the B200 is still the parity gate, and performance / memory-model behaviour are out of its
reach.  The product never loads the emulated library: this file passes its path to a pytest
subprocess through the loader's BENDY2D_B200_LIB override.
"""
import os
import shutil
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "cuemu"))

pytestmark = pytest.mark.skipif(shutil.which(os.environ.get("CXX", "g++")) is None, reason="needs g++")


@pytest.fixture(scope="module")
def emu_lib():
    import build as cuemu_build

    return cuemu_build.build()


def run_gpu_tests_on_emulator(emu_lib, args, timeout=900, extra_env=None):
    env = dict(os.environ)
    env.update({"BENDY2D_B200_LIB": emu_lib, "BENDY_CUDA_EMU": "1"})
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] + args
    try:  # the emulated suites are CPU-bound and independent: spread them over a few cores when xdist is there
        import xdist  # noqa: F401

        cmd += ["-n", str(max(1, min(4, (os.cpu_count() or 2) // 2)))]
    except ImportError:
        pass
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-40:])
    assert r.returncode == 0, f"emulated run failed ({' '.join(args)}):\n{tail}"
    assert " passed" in r.stdout, tail
    return r.stdout


def test_emulator_primitives_self_check(emu_lib, tmp_path):
    """the emulation itself: warp primitives, barriers, per-CTA shared memory, grid batches, canaries"""
    exe = tmp_path / "selfcheck"
    subprocess.check_call([os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-w", "-I", os.path.join(HERE, "cuemu", "include"),
                           os.path.join(HERE, "cuemu", "selfcheck.cpp"), os.path.join(HERE, "cuemu", "runtime.cpp"),
                           "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "selfcheck ok" in r.stdout
    # an out-of-bounds device write and a barrier that not every thread reaches must both be fatal
    for mode, msg in (("oob", "out-of-bounds write"), ("deadlock", "deadlock")):
        r = subprocess.run([str(exe), mode], capture_output=True, text=True, timeout=120)
        assert r.returncode != 0 and msg in r.stderr, (mode, r.stdout, r.stderr)


def test_counting_sort_with_the_one_pass_scan_at_kernel_level(emu_lib, tmp_path):
    """kernel-level (tests/cuemu/kernel_unit.cpp): k2_count (histogram + scan-tile totals, CTA tables overflowing),
    k2_scan and k2_scatter checked cell by cell, on a grid smaller and one larger than the emulated device (the
    scan has no inter-CTA barrier any more, so nothing depends on residency)"""
    build_dir = os.path.join(HERE, "cuemu", "_build")
    exe = tmp_path / "kernel_unit"
    subprocess.check_call([os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-w", "-ffp-contract=off",
                           "-DBENDY_SPIN_HOOK()=cuemu::yield_spin()", "-DBENDY_SPIN_LIMIT=150000000ll",
                           "-I", os.path.join(HERE, "cuemu", "include"), "-I", os.path.join(build_dir, "bendy2d_b200", "csrc"),
                           os.path.join(HERE, "cuemu", "kernel_unit.cpp"), os.path.join(HERE, "cuemu", "runtime.cpp"),
                           "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "kernel_unit ok" in r.stdout, r.stdout + r.stderr


def test_parity_suite_on_emulator(emu_lib):
    out = run_gpu_tests_on_emulator(emu_lib, ["tests/test_gpu_parity.py", "-k", "not full_size"])
    assert "failed" not in out


def test_parity_suite_on_emulator_plain_launches(emu_lib):
    """same kernels without the small-scene single launch and without programmatic dependent launch"""
    run_gpu_tests_on_emulator(emu_lib, ["tests/test_gpu_parity.py", "-k", "c1_reference or c2_small or c3_small or c4_small"],
                              extra_env={"BENDY_SMALL_SCENE": "0", "BENDY_PDL": "0"})


@pytest.mark.parametrize("order", ["2"])
def test_results_do_not_depend_on_the_thread_schedule(emu_lib, order):
    """A device promises no execution order.  The emulation resumes runnable threads last-first (1) or in a new
    pseudo-random order every sweep (2): every bit-exact parity test must still pass, i.e. slot orders, atomic
    arrival orders and near-list orders never leak into results."""
    run_gpu_tests_on_emulator(emu_lib, ["tests/test_gpu_parity.py", "tests/test_z_gpu_switches.py", "tests/test_golden.py",
                                        "tests/test_z_gpu_strips_replicated.py", "-k", "not full_size"],
                              extra_env={"CUEMU_ORDER": order})


def test_limits_strips_snapshot_golden_on_emulator(emu_lib):
    run_gpu_tests_on_emulator(emu_lib, ["tests/test_gpu_limits.py", "tests/test_gpu_strips.py", "tests/test_snapshot.py",
                                        "tests/test_golden.py", "tests/test_z_gpu_strips_replicated.py", "-k", "not 4200 and not nccl"])


def test_scan_shapes_and_switches_on_emulator(emu_lib):
    """tests/test_z_gpu_switches.py: the one-pass scan over grid shapes, scheduling switches"""
    run_gpu_tests_on_emulator(emu_lib, ["tests/test_z_gpu_switches.py"])


def test_randomised_worlds_on_emulator(emu_lib):
    """tests/test_z_gpu_fuzz.py: seeded random worlds mixing every feature, bit-exact where reference-pinned"""
    run_gpu_tests_on_emulator(emu_lib, ["tests/test_z_gpu_fuzz.py"], extra_env={"BENDY_FUZZ_SEEDS": "60"})
