"""Single-process multi-GPU context (tritd_create_devices) -- the mode a MEX gateway uses from one MATLAB process:
full tensors in and out, slabs and peer mailboxes inside.  Needs >= 2 GPUs with peer access (gpurun --gpus 2);
skipped on a one-GPU box."""
import numpy as np
import pytest

import tritd
import tritd_oracle as orc
from conftest import rel_err
from tritd import synth

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")


@needs2
@pytest.mark.parametrize("cfg,shape,iters", [("cfg1", (40, 36, 25), 8), ("cfg3", (48, 64, 21), 6), ("cfg5", (72, 40, 9), 5)])
def test_devices_context_equals_single_gpu_and_oracle(cfg, shape, iters):
    nd = min(_ngpu(), 4)
    w = synth.make_config(cfg, shrink=shape)
    o = dict(w["opts"], maxIter=iters, tol=0.0)
    ref = orc.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"])
    one = tritd.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"])
    with tritd.Context.from_devices(list(range(nd))) as g:
        for _ in range(2):                                       # second call: cached device state is reused
            A, B, C, O, eh, info = tritd.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"], ctx=g, return_info=True,
                                                            want_L=True, want_E=True)
            for x, y, z in zip((A, B, C, O, eh), ref, one):
                assert rel_err(x, y) < 1e-8 and rel_err(x, z) < 1e-10
            assert rel_err(info["L"], orc.triple_product(*ref[:3])) < 1e-8
            assert info["launches"] > 0


@needs2
def test_devices_context_masked_and_stop_rule():
    shape, r = (40, 36, 24), 3
    D = synth.make_lowrank_sparse(*shape, r, 0.05, 21)
    F = synth.init_factors(*shape, r, 22)
    m = np.random.default_rng(23).random(shape) >= 0.3
    o = dict(synth.TRAFFIC_OPTS, maxIter=25, tol=0.0)
    ref = orc.triple_ADMM_masked(D, m, r, o, *F)
    with tritd.Context.from_devices([0, 1]) as g:
        A, B, C, O, E, out = tritd.triple_ADMM_masked(D, m, r, o, *F, ctx=g)
        assert rel_err(A, ref[0]) < 1e-8 and rel_err(C, ref[2]) < 1e-8 and rel_err(E, ref[4]) < 1e-8
        # the golden "stop" case: the relative-change rule fires at the same iteration on every device
        import make_golden
        Ds, rs, os_, a0, b0, c0 = make_golden.case_inputs("stop_30x30x30_r3")
        res = tritd.triple_decomp_ADMM(Ds, rs, os_, a0, b0, c0, ctx=g)
        refs = orc.triple_decomp_ADMM(Ds, rs, os_, a0, b0, c0)
        assert len(res[4]) == len(refs[4]) < os_["maxIter"]
        assert rel_err(res[0], refs[0]) < 1e-8 and rel_err(res[3], refs[3]) < 1e-8


def test_devices_context_with_one_device_is_a_plain_context():
    with tritd.Context.from_devices([0]) as g:
        w = synth.make_config("cfg1", shrink=(20, 18, 12))
        o = dict(w["opts"], maxIter=4, tol=0.0)
        a = tritd.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"], ctx=g)
        b = tritd.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"])
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    with pytest.raises(tritd.TritdError):
        tritd.Context.from_devices([0, 0])
