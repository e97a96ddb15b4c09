"""Slab mode (BASELINE config 5): ONE domain advanced by several devices of one box through a shared address range
(rlfc_config.n_devices > 1).  The kernels, their arithmetic and its order are the single-device ones, so the results must
be bit-identical to the single-device run and to the oracle.  Needs >= 2 GPUs (skipped otherwise): run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_slab.py -m gpu`."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def device_count():
    try:
        rt = C.CDLL("libcudart.so")
    except OSError:
        try:
            import torch
            return torch.cuda.device_count()
        except Exception:
            return 0
    n = C.c_int(0)
    return n.value if rt.cudaGetDeviceCount(C.byref(n)) == 0 else 0


def assert_same(a, b, what):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    assert not bad.any(), f"{what}: {int(bad.sum())} of {a.size} values differ (max |diff| {np.nanmax(np.abs(a - b)):.3e})"


def run(rlfc, n_devices, resolution, steps, act_at=1):
    t_step = float(np.float32(0.18) / np.float32(resolution))
    out = []
    with rlfc.AFCCylinderBatch(1, init_state=None, resolution=resolution, x_lengths=16, y_lengths=8, t_step=t_step,
                               n_devices=n_devices) as env:
        for k in range(steps):
            f = env.update2(np.array([[0.5, -0.5]], np.float32) if k == act_at else None)
            out.append((f[0].copy(), tuple(env.mg_iters()[0])))
        fields = env.get_fields(0)
    return out, fields


@pytest.mark.skipif(device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("resolution,n_devices", [(64, 2), (128, 2)])
def test_slab_equals_single_device(rlfc, resolution, n_devices):
    """1024x512 (fields smaller than a stripe: compute split only) and 2048x1024 (fields striped over the devices):
    forces, MG iteration counts and all fields after 3 steps == the single-device run."""
    ref, ref_fields = run(rlfc, 1, resolution, 3)
    got, got_fields = run(rlfc, n_devices, resolution, 3)
    for k, ((fr, ir), (fg, ig)) in enumerate(zip(ref, got)):
        assert_same(fg, fr, f"force step {k}")
        assert ig == ir, f"MG iterations step {k}: {ig} != {ir}"
    for nm, a, b in zip(("ux", "uy", "p"), got_fields, ref_fields):
        assert_same(a, b, nm)


@pytest.mark.skipif(device_count() < 2, reason="needs 2 GPUs")
def test_slab_equals_oracle(rlfc, oracle):
    """The 1024x512 domain on 2 devices against the CPU oracle (impulsive start, a non-zero action): every float equal."""
    resolution = 64
    t_step = np.float32(0.18) / np.float32(resolution)
    ref = oracle.OracleEnv(literal=False, resolution=resolution, xLengths=16, yLengths=8, tStep=float(t_step))
    ref.set_xi(0.5, -0.5)
    with rlfc.AFCCylinderBatch(1, init_state=None, resolution=resolution, x_lengths=16, y_lengths=8, t_step=float(t_step),
                               n_devices=2) as env:
        for k in range(2):
            f = env.update2(np.array([[0.5, -0.5]], np.float32) if k == 0 else None)
            ref.update2()
            assert_same(f[0], np.array(ref.force(), np.float32), f"force step {k}")
            assert tuple(env.mg_iters()[0]) == ref.mg_iters()
        for nm, a, b in zip(("ux", "uy", "p"), env.get_fields(0), ref.get_state()):
            assert_same(a, b, nm)


@pytest.mark.skipif(device_count() < 4, reason="needs 4 GPUs")
def test_slab_four_devices(rlfc):
    ref, ref_fields = run(rlfc, 1, 128, 2, act_at=0)
    got, got_fields = run(rlfc, 4, 128, 2, act_at=0)
    for k, ((fr, ir), (fg, ig)) in enumerate(zip(ref, got)):
        assert_same(fg, fr, f"force step {k}")
        assert ig == ir
    for nm, a, b in zip(("ux", "uy", "p"), got_fields, ref_fields):
        assert_same(a, b, nm)
