"""Tensor-parallel algebra on the CPU, two gloo ranks: "quantize first, then shard the integers" makes
the sharded GEMVs add up EXACTLY to the unsharded one in the integer domain -- column-parallel shards concatenate,
row-parallel shards sum -- which is what lets the GPU engine's TP tokens equal its single-GPU tokens.  Also covers the
rank plumbing the bench uses (object broadcast of the group id, max-over-ranks reduction)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_no, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        port = oracle.port()
        rng = np.random.default_rng(11)
        K, N = 256, 96
        w = rng.uniform(-0.07, 0.07, (K, N)).astype(np.float32)
        xf = rng.integers(-(2 ** 22), 2 ** 22, size=K).astype(np.int64)      # the fixed-point activations of gemv.cuh
        ok = True
        for qt in (oracle.QINT4, oracle.QINT8):
            s, z = port.quant_info(w, qt, True)                              # parameters of the WHOLE tensor on every rank
            q = port.quantize(w, qt, s, z).astype(np.int64)
            full = xf @ q
            # row-parallel (o / down): each rank owns K/world rows, partial sums are all-reduced
            rows = slice(rank * K // world, (rank + 1) * K // world)
            part = torch.from_numpy(xf[rows] @ q[rows])
            dist.all_reduce(part, op=dist.ReduceOp.SUM)
            ok &= bool(np.array_equal(part.numpy(), full))
            # column-parallel (q/k/v, gate/up): each rank owns N/world columns, results are concatenated
            cols = slice(rank * N // world, (rank + 1) * N // world)
            mine = torch.from_numpy(np.ascontiguousarray(xf @ q[:, cols]))
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            ok &= bool(np.array_equal(torch.cat(parts).numpy(), full))
            # a shard quantized on its own would NOT be the reference's integers (different min/max -> different scale)
            s_shard, _ = port.quant_info(np.ascontiguousarray(w[rows] * (0.5 if rank else 1.0)), qt, True)
            ok &= (s_shard != s) or rank == 0
        # The exchange of the fused engine: every rank hands its fp32 partial of a row-parallel GEMV to every rank as words tagged
        # with the exchange number, and every rank adds residual + partial_0 + ... + partial_{P-1} IN RANK ORDER from what was
        # stored -- the same additions on every rank, so the replicated residual stream stays bit-identical (a tree or
        # arrival-order sum would not guarantee that).  Restated with an all_gather of {value bits, tag} words.
        resid = rng.standard_normal(N).astype(np.float32)
        for seq in (1, 2, 3):
            partial = np.random.default_rng(100 * seq + rank).standard_normal(N).astype(np.float32)
            words = torch.from_numpy(((np.uint64(seq) << np.uint64(32)) | partial.view(np.uint32).astype(np.uint64)).view(np.int64))
            got = [torch.empty_like(words) for _ in range(world)]
            dist.all_gather(got, words)
            acc = resid.copy()
            for r in range(world):
                w64 = got[r].numpy().view(np.uint64)
                ok &= bool(np.all((w64 >> np.uint64(32)) == np.uint64(seq)))            # every word carries this exchange's number
                acc = acc + (w64 & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.float32)
            mine = torch.from_numpy(acc.view(np.int32).copy())
            everyone = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(everyone, mine)
            ok &= all(torch.equal(everyone[0], e) for e in everyone)                   # bit-identical on every rank
            expect = resid.copy()
            for r in range(world):
                expect = expect + np.random.default_rng(100 * seq + r).standard_normal(N).astype(np.float32)
            ok &= bool(np.array_equal(acc, expect))
            resid = acc
        # rank plumbing of bench.py: id broadcast over a side channel, max over ranks of the timed region
        box = [b"\\x07" * 128 if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ok &= box[0] == b"\\x07" * 128
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok &= float(t[0]) == 10.0 + world - 1
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_tp_integer_algebra_two_gloo_ranks():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
