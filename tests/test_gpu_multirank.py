"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): two ranks, one process per GPU, mode-3 slabs,
NCCL all-reduce of [RHS_A ; C3'C3], RHS_B and the residual norms; results must equal the single-GPU
solve to 1e-10 and the oracle to 1e-8, uneven slab split included."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, r, iters, out_dir, nccl_path):
    if nccl_path:
        os.environ["TRITD_XCHG_NCCL"] = "1"        # NCCL all-reduces instead of the NVLink peer mailboxes
    for p in (os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"),):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import tritd
    from tritd import synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    ids = [tritd.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = tritd.Context(rank, rank, world, ids[0])
    w = synth.make_config("cfg1", shrink=shape)
    w["A0"], w["B0"], w["C0"] = synth.init_factors(*shape, r, 77)
    o = dict(w["opts"], maxIter=iters, tol=0.0)
    t0, t1 = tritd.slab_bounds(shape[2], world, rank)
    Ds = np.asfortranarray(w["D"][:, :, t0:t1]); C0s = np.asfortranarray(w["C0"][:, :, t0:t1])
    A, B, C, O, eh = tritd.triple_decomp_ADMM(Ds, r, o, w["A0"], w["B0"], C0s, ctx=ctx)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), A=A, B=B, C=C, O=O, eh=eh, t0=t0, t1=t1)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("nccl_path", [False, True], ids=["peer_mailboxes", "nccl_allreduce"])
@pytest.mark.parametrize("shape,r,iters", [((40, 36, 25), 5, 8), ((130, 70, 11), 4, 5), ((96, 80, 9), 8, 4)])
def test_two_gpus_equal_one_gpu_and_oracle(shape, r, iters, nccl_path, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import tritd
    import tritd_oracle as orc
    from conftest import rel_err
    from tritd import synth

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), shape, r, iters, str(tmp_path), nccl_path), nprocs=world, join=True)
    w = synth.make_config("cfg1", shrink=shape)
    w["A0"], w["B0"], w["C0"] = synth.init_factors(*shape, r, 77)
    o = dict(w["opts"], maxIter=iters, tol=0.0)
    one = tritd.triple_decomp_ADMM(w["D"], r, o, w["A0"], w["B0"], w["C0"])
    ref = orc.triple_decomp_ADMM(w["D"], r, o, w["A0"], w["B0"], w["C0"])
    parts = [np.load(os.path.join(str(tmp_path), f"rank{g}.npz")) for g in range(world)]
    C = np.concatenate([p["C"] for p in parts], axis=2)
    O = np.concatenate([p["O"] for p in parts], axis=2)
    for p in parts:
        assert rel_err(p["eh"], one[4]) < 1e-10 and rel_err(p["A"], one[0]) < 1e-10 and rel_err(p["B"], one[1]) < 1e-10
    assert np.array_equal(parts[0]["A"], parts[1]["A"]) and np.array_equal(parts[0]["eh"], parts[1]["eh"])   # replicas agree bitwise
    assert rel_err(C, one[2]) < 1e-10 and rel_err(O, one[3]) < 1e-10
    assert rel_err(C, ref[2]) < 1e-8 and rel_err(O, ref[3]) < 1e-8 and rel_err(parts[0]["eh"], ref[4]) < 1e-8
