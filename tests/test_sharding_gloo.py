"""N>1 host-side logic on CPU: world_size-2 (and 3, 4) gloo processes run the mode-3 sharded restatement of the
iteration (oracle/tritd_oracle_sharded.py), exchanging [RHS_A ; C3'C3], RHS_B and the residual norms
exactly where libtritd does, and must reproduce the unsharded oracle -- including an uneven split.
Two exchange models: a gloo all-reduce (libtritd's NCCL path) and the peer-mailbox rule of
csrc/kernels_xchg.cuh (every rank receives all ranks' partials and sums them in RANK ORDER), under which
the replicated factors and the error history must be BITWISE equal on all ranks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, n3, iters, out_dir, exchange):
    for p in (os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"), os.path.join(ROOT, "oracle"),
              os.path.join(ROOT, "tests", "golden")):
        sys.path.insert(0, p)
    import make_golden
    import tritd_oracle as orc
    import tritd_oracle_sharded as orcs
    from tritd import synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D, r, o, A0, B0, C0 = make_golden.case_inputs(case)
    D = np.asfortranarray(D[:, :, :n3]); C0 = np.asfortranarray(C0[:, :, :n3])
    o = dict(o, maxIter=iters, tol=0.0)
    t0, t1 = synth.slab_bounds(n3, world)[rank]

    epoch = [1]                                                      # (0 = the zeroed mailbox: never a valid epoch)

    def allreduce(x):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
        if exchange == "allreduce":
            dist.all_reduce(t)
            return t.numpy()
        # the mailbox of csrc/kernels_xchg.cuh: one slot per source rank, every double travelling as the self-validating
        # 16-byte word {lo, epoch, hi, epoch}; the receiver accepts a word only when both halves show the epoch of this
        # exchange, then sums the ranks' values in rank order
        epoch[0] += 1
        bits = t.numpy().view(np.uint64).ravel()
        words = np.empty((bits.size, 4), dtype=np.uint32)
        words[:, 0] = (bits & 0xFFFFFFFF).astype(np.uint32); words[:, 2] = (bits >> 32).astype(np.uint32)
        words[:, 1] = words[:, 3] = epoch[0]
        w = torch.from_numpy(words.view(np.int32).copy())
        slots = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(slots, w)
        acc = None
        for s_ in slots:                                             # rank order, the same on every rank
            q = s_.numpy().view(np.uint32)
            assert (q[:, 1] == epoch[0]).all() and (q[:, 3] == epoch[0]).all()
            stale = q.copy(); stale[:, 3] -= 1                       # a half from the previous exchange must not validate
            assert not ((stale[:, 1] == epoch[0]) & (stale[:, 3] == epoch[0])).any()
            v = ((q[:, 2].astype(np.uint64) << np.uint64(32)) | q[:, 0].astype(np.uint64)).view(np.float64).reshape(t.shape)
            acc = v.copy() if acc is None else acc + v
        return acc

    A1, B2, C3 = orc.factors_to_unfolded(A0, B0, C0)
    A1s, B2s, C3s, Os, ehs = orcs.admm_sharded(np.asfortranarray(D[:, :, t0:t1]), r, o, A1, B2, C3[t0:t1], allreduce)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), A1=A1s, B2=B2s, C3=C3s, O=Os, eh=ehs, t0=t0, t1=t1)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["allreduce", "mailbox"])
@pytest.mark.parametrize("world,case,n3,iters", [(2, "small_40x36x24_r5", 24, 6), (2, "odd_33x17x9_r2", 9, 6),
                                                 (3, "small_40x36x24_r5", 23, 4), (4, "odd_33x17x9_r2", 9, 4)])
def test_sharded_iteration_equals_oracle(world, case, n3, iters, exchange, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    import tritd_oracle as orc
    from conftest import rel_err
    from tritd import synth

    mp.spawn(_worker, args=(world, _free_port(), case, n3, iters, str(tmp_path), exchange), nprocs=world, join=True)
    D, r, o, A0, B0, C0 = make_golden.case_inputs(case)
    D = np.asfortranarray(D[:, :, :n3]); C0 = np.asfortranarray(C0[:, :, :n3])
    o = dict(o, maxIter=iters, tol=0.0)
    A, B, C, O, eh = orc.triple_decomp_ADMM(D, r, o, A0, B0, C0)
    parts = [np.load(os.path.join(str(tmp_path), f"rank{g}.npz")) for g in range(world)]
    # contiguous slabs, the first n3 % world ranks one slice longer (tritd_slab_bounds): 12+12, 5+4, 8+8+7, 3+2+2+2
    sizes = [int(p["t1"] - p["t0"]) for p in parts]
    assert sizes == {(2, 24): [12, 12], (2, 9): [5, 4], (3, 23): [8, 8, 7], (4, 9): [3, 2, 2, 2]}[(world, n3)]
    assert [int(p["t0"]) for p in parts] == [sum(sizes[:g]) for g in range(world)]
    assert [tuple(b) for b in synth.slab_bounds(n3, world)] == [(int(p["t0"]), int(p["t1"])) for p in parts]
    for p in parts:
        assert rel_err(p["eh"], eh) < 1e-10                                                 # identical on every rank
        assert rel_err(p["A1"], orc.unfold(A, 1)) < 1e-9 and rel_err(p["B2"], orc.unfold(B, 2)) < 1e-9   # replicated
    if exchange == "mailbox":          # rank-ordered sums: replicas are bitwise equal, no broadcast needed
        for p in parts[1:]:
            assert np.array_equal(parts[0]["A1"], p["A1"]) and np.array_equal(parts[0]["B2"], p["B2"])
            assert np.array_equal(parts[0]["eh"], p["eh"])
    C3 = np.concatenate([p["C3"] for p in parts], axis=0)
    Ocat = np.concatenate([p["O"] for p in parts], axis=2)
    assert rel_err(C3, orc.unfold(C, 3)) < 1e-9 and rel_err(Ocat, O) < 1e-9                  # slabs tile the tensor
