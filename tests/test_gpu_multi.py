"""Multi-GPU fan-out inside the library: petlion(..., devices=[...]) shards every batch
contiguously over the GPUs of the box, each shard runs on its own device, and one ncclAllGather leaves the whole
batch's 80-byte summaries on every device.  Needs two GPUs (the driver's single-GPU test box skips it; it is run by
hand under `gpurun --gpus 2`, log in profiles/)."""
import ctypes as C

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _two_gpus():
    import torch
    return torch.cuda.device_count() >= 2


@pytest.mark.parametrize("B", [37, 64])
def test_group_equals_single_device(B):
    if not _two_gpus():
        pytest.skip("needs two GPUs")
    import torch
    import petlion_b200 as P
    from petlion_b200 import _lib, sweep
    one = P.petlion("LCO")
    two = P.petlion("LCO", devices=[0, 1])
    assert two.θ_keys == one.θ_keys and two.N.tot == one.N.tot
    th = sweep.randomised_theta(one, B, first=4000)
    util.set_theta_batch(one, th); util.set_theta_batch(two, th)
    a = P.simulate(one, I=2, SOC=0, V_max=4.1)
    b = P.simulate(two, I=2, SOC=0, V_max=4.1)
    # a system's bits depend neither on the device nor on the shard it lands in
    assert a.results[-1].summary.tobytes() == b.results[-1].summary.tobytes()
    assert np.array_equal(a.Y, b.Y) and np.array_equal(a.V, b.V, equal_nan=True) and np.array_equal(a.n_points, b.n_points)
    # simulate!(): the host state arrays are the hand-off, shard by shard
    P.simulate_(a, one, V="hold", V_max=4.1, I_min=0.05)
    P.simulate_(b, two, V="hold", V_max=4.1, I_min=0.05)
    assert a.results[-1].summary.tobytes() == b.results[-1].summary.tobytes()
    # the one collective: every device holds the whole batch's summaries
    R = (B + 1) // 2
    host = b.results[-1].summary
    for k in (0, 1):
        ptr, rows = two.device_summaries(k)
        assert rows == R
        with torch.cuda.device(k):
            buf = torch.empty(2 * R * 10, dtype=torch.float64, device=f"cuda:{k}")
            rt = C.CDLL("libcudart.so.12")                  # (loaded by torch already)
            rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
            assert rt.cudaMemcpy(buf.data_ptr(), ptr, 2 * R * 80, 3) == 0      # cudaMemcpyDeviceToDevice
            got = buf.cpu().numpy().view(_lib.SUMMARY_DTYPE).reshape(-1)
        assert got[:B].tobytes() == host.tobytes()
        assert not got[B:].tobytes().strip(b"\0")            # the tail rows of the last block are zero
    assert _lib.lib().plb_group_last_gather_ms(two._g) > 0.0
    with pytest.raises(NotImplementedError):
        P.simulate(two, np.array([0.0, 10.0, 1e6]), I=-1, SOC=1)


def test_group_of_one_device_is_a_plain_model():
    import petlion_b200 as P
    p = P.petlion("LCO", devices=[0])
    assert p._g is None
    assert P.simulate(p, I=-1, SOC=1).results[-1].summary["n_steps"][0] == 80
