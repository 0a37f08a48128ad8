"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/allwave_cuda.h
declares, refuses to compute without a GPU (no CPU fallback), and the host-side partitioner shards
a pair list correctly (single process and world_size-2 gloo)."""
import os
import re
import socket

import pytest

import allwave_b200 as aw
from allwave_b200 import partition, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "allwave_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(aw_[a-z0-9_]+)\s*\(", hdr)) - {"aw_result_cb"}
    assert declared == set(aw._cabi.EXPORTS), declared ^ set(aw._cabi.EXPORTS)
    L = aw._cabi.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.aw_abi_version() == 1
    assert L.aw_strerror(aw._cabi.AW_ENODEVICE) == b"no usable CUDA device"


def test_no_cpu_fallback_without_device():
    L = aw._cabi.lib()
    if L.aw_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(aw.AllwaveError) as e:
        aw.Context(0)
    assert e.value.status == aw._cabi.AW_ENODEVICE and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "allwave_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".c", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in src and "aw_oracle.h" not in src and "liballwave_oracle" not in src, f


def test_synth_deterministic_and_shapes():
    a = synth.generate(2, 5, 3000, 0.05)
    b = synth.generate(2, 5, 3000, 0.05)
    assert a == b and all(set(s) <= set(b"ACGT") for s in a[1])
    assert synth.generate(3, 5, 3000, 0.05)[1] != a[1]
    _, seqs, rc = synth.generate(5, 40, 500, 0.03, rc_prob=0.5)
    assert 5 < sum(rc) < 35
    _, sv, _ = synth.generate(4, 2, 20000, 0.001, 0.02, n_sv=5)
    assert all(abs(len(s) - 20000) < 8000 for s in sv)


def test_partition_lpt():
    lens = [1000 + 37 * i for i in range(30)]
    pairs = [(i, j) for i in range(30) for j in range(30) if i != j]
    for n in (1, 2, 4, 8):
        shards = partition.partition_pairs(pairs, lens, n)
        assert len(shards) == n and sorted(p for s in shards for p in s) == sorted(pairs)
        loads = partition.shard_loads(shards, lens)
        assert max(loads) / (sum(loads) / n) < 1.02
    # heavy-tailed costs: LPT keeps the max load near the mean
    lens = [100] * 20 + [100000] * 3
    pairs = [(i, j) for i in range(23) for j in range(23) if i != j]
    loads = partition.shard_loads(partition.partition_pairs(pairs, lens, 4), lens)
    assert max(loads) / (sum(loads) / 4) < 1.3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = [500 + 11 * i for i in range(16)]
    pairs = [(i, j) for i in range(16) for j in range(16) if i != j]
    mine = partition.partition_pairs(pairs, lens, world)[rank]
    # "results are gathered": every rank contributes its pair count and a checksum, max-over-ranks timing
    t = torch.tensor([len(mine), sum(a * 131 + b for a, b in mine)], dtype=torch.int64)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    q.put((rank, [g.tolist() for g in gathered], float(ms[0]), mine))
    dist.destroy_process_group()


def test_partition_world2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    out.sort()
    all_pairs = sorted(out[0][3] + out[1][3])
    assert all_pairs == sorted((i, j) for i in range(16) for j in range(16) if i != j)
    assert out[0][1] == out[1][1] and sum(g[0] for g in out[0][1]) == 240
    assert out[0][2] == out[1][2] == 11.0
