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


# ---- C++ host mirror (allwave_b200/host/allwave.hpp) against the oracle ----
def test_host_pair_lists_match_oracle(oracle):
    from allwave_b200 import hostlib as H

    ids = ["s%06d" % i for i in range(50)]
    assert H.pair_list(ids[:3]) == oracle.pair_list(ids[:3], None) and len(H.pair_list(ids[:6])) == 30
    assert H.pair_list(ids[:4], exclude_self=False) == oracle.pair_list(ids[:4], None, exclude_self=False)
    assert H.pair_list(ids, H.KIND_RANDOM, 0.3) == oracle.pair_list(ids, None, kind=oracle.SPARS_RANDOM, fraction=0.3)
    assert H.pair_list(ids, H.KIND_CONNECTIVITY, 0.99) == oracle.pair_list(ids, None, kind=oracle.SPARS_GIANT, fraction=0.99)
    assert H.pair_list(ids, H.KIND_AUTO) == oracle.pair_list(ids, None, kind=oracle.SPARS_AUTO)
    L = H.lib()
    for n in (1, 2, 3, 5, 10, 11, 200, 5000):
        assert L.awh_connectivity_probability(n, 0.99) == oracle.lib().awo_connectivity_probability(n, 0.99)
    for s in ("seq1:seq2", "s000000:s000001", ""):
        assert L.awh_hash_str(s.encode()) == oracle.hash_str(s)
    m = [[0.0, 0.1, 0.5], [0.1, 0.0, 0.3], [0.5, 0.3, 0.0]]
    for far in (False, True):
        assert H.build_knn_graph(m, 1, far) == oracle.build_knn_graph(m, 1, far)


def test_host_parsers():
    from allwave_b200 import hostlib as H

    p = H.parse_scores("0,5,8,2,24,1")
    assert (p.mismatch_penalty, p.gap_open, p.gap_extend, p.gap2_open, p.gap2_extend, p.has_gap2_open) == (5, 8, 2, 24, 1, 1)
    assert H.lib().awh_mode_from_params(p) == 2 and H.lib().awh_mode_from_params(H.parse_scores("0,1,1,1")) == 0
    assert H.lib().awh_mode_from_params(H.parse_scores("0,4,6,2")) == 1
    with pytest.raises(ValueError, match="Invalid number of scores: 3. Expected 4 or 6 values."):
        H.parse_scores("0,1,1")
    with pytest.raises(ValueError, match="Failed to parse scores"):
        H.parse_scores("0,x,1,1")
    # the -p grammar and its messages (src/main.rs:136-203)
    assert H.parse_sparsification("none")["kind"] == H.KIND_NONE and H.parse_sparsification("auto")["kind"] == H.KIND_AUTO
    assert H.parse_sparsification("giant:0.99") == dict(kind=H.KIND_CONNECTIVITY, value=0.99, k_nearest=0, k_farthest=0, random_fraction=0.0, kmer_size=0)
    assert H.parse_sparsification("connectivity:0.5")["value"] == 0.5
    t = H.parse_sparsification("tree:2:1:0.1")
    assert (t["kind"], t["k_nearest"], t["k_farthest"], t["random_fraction"], t["kmer_size"]) == (H.KIND_TREE, 2, 1, 0.1, 0)
    assert H.parse_sparsification("tree:3:0:0:21")["kmer_size"] == 21
    for bad, msg in (("random:0", "Random fraction must be between 0 and 1"), ("random:x", "Invalid random fraction"),
                     ("giant:1.0", "Giant component probability must be between 0 and 1"), ("tree:0:0:0.1", "At least one of k_nearest"),
                     ("tree:1:1", "Invalid tree format"), ("tree:1:1:0.1:2", "K-mer size must be between 3 and 31"),
                     ("tree:1:1:1.5", "Random fraction must be between 0 and 1"), ("bogus", "Invalid sparsification strategy")):
        with pytest.raises(ValueError, match=msg):
            H.parse_sparsification(bad)


def test_cpp_partition_matches_python():
    """the C++ host's multi-GPU sharding (allwave.hpp partition_pairs) is the same LPT assignment as partition.py"""
    import random

    from allwave_b200 import hostlib as H
    from allwave_b200 import partition

    rnd = random.Random(11)
    lens = [rnd.choice([150, 5000, 10000, 10000, 250000]) + rnd.randrange(50) for _ in range(40)]
    pairs = [(i, j) for i in range(40) for j in range(40) if i != j and rnd.random() < 0.3]
    for parts in (1, 2, 3, 8):
        got = H.partition_pairs(pairs, lens, parts)
        exp = partition.partition_pairs(pairs, lens, parts)
        assert got == exp
        assert sorted(p for s in got for p in s) == sorted(pairs)
        if parts > 1:
            loads = partition.shard_loads(got, lens)
            assert max(loads) <= 1.25 * (sum(loads) / parts) + max(partition.predicted_cost(lens[q], lens[t]) for q, t in pairs)


def test_cli_surface_presets_filters_gz(tmp_path):
    """the CLI rows of SURVEY 8(f): -x ANI presets (src/main.rs:83-124), -k / -e prefix filters (:237-277), .gz FASTA input"""
    import gzip

    from allwave_b200 import hostlib as H

    for s, exp in [("98%", "0,7,12,2,36,1"), ("95", "0,7,12,2,36,1"), ("0.9", "0,5,8,2,24,1"), ("85%", "0,5,8,2,24,1"), ("80", "0,4,6,2,18,1"),
                   ("0.7", "0,3,4,1"), ("60%", "0,1,1,1"), ("1.0", "0,7,12,2,36,1")]:
        assert H.parse_ani_preset(s) == exp
    for bad in ("40%", "101", "1.5", "abc", "0.0", ""):
        with pytest.raises(ValueError):
            H.parse_ani_preset(bad)
    recs = [("sampleA#1#chr1 desc", "ACGTACGTAC"), ("sampleA#2#chr1", "ACGT\nACGT"), ("sampleB#1#chr1", "GGGG"), ("other", "TT")]
    text = "".join(f">{i}\n{s}\n" for i, s in recs)
    fa = tmp_path / "x.fa"
    fa.write_text(text)
    ids, tot = H.read_fasta(fa)
    assert ids == ["sampleA#1#chr1", "sampleA#2#chr1", "sampleB#1#chr1", "other"] and tot == 10 + 8 + 4 + 2
    assert H.read_fasta(fa, keep_prefixes="sampleA, other")[0] == ["sampleA#1#chr1", "sampleA#2#chr1", "other"]
    assert H.read_fasta(fa, exclude_prefixes="sampleA#1,sampleB")[0] == ["sampleA#2#chr1", "other"]
    gz = tmp_path / "x.fa.gz"
    with gzip.open(gz, "wb") as f:  # two gzip members back to back, like bgzf blocks
        f.write(text[:30].encode())
    with open(gz, "ab") as f:
        f.write(gzip.compress(text[30:].encode()))
    assert H.read_fasta(gz) == (ids, tot)


def _c_type_to_rust(t):
    """canonical Rust spelling of a C parameter / return type of include/allwave_cuda.h"""
    t = re.sub(r"\s+", " ", t.strip())
    t = re.sub(r"\s*\*\s*", "*", t)
    table = {
        "void": "()", "int": "c_int", "int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "uint8_t": "u8", "float": "f32",
        "const char*": "*const c_char", "char*": "*mut c_char", "void*": "*mut c_void", "const uint8_t*": "*const u8", "uint8_t*": "*mut u8",
        "const uint64_t*": "*const u64", "uint64_t*": "*mut u64", "uint32_t*": "*mut u32", "float*": "*mut f32",
        "const uint8_t*const*": "*const *const u8", "const char*const*": "*const *const c_char",
        "aw_ctx*": "*mut aw_ctx", "const aw_ctx*": "*const aw_ctx", "aw_ctx**": "*mut *mut aw_ctx",
        "aw_batch*": "*mut aw_batch", "aw_batch**": "*mut *mut aw_batch",
        "aw_aligner*": "*mut aw_aligner", "const aw_aligner*": "*const aw_aligner", "aw_aligner**": "*mut *mut aw_aligner",
        "const aw_params*": "*const aw_params", "const aw_pair*": "*const aw_pair",
        "aw_result_cb": "aw_result_cb", "aw_chunk_source": "aw_chunk_source", "aw_paf_block_cb": "aw_paf_block_cb",
    }
    return table[t]


def test_rust_sys_crate_matches_header():
    """bindings/allwave-cuda-sys/src/lib.rs (source only: no Rust toolchain in this image) declares exactly the functions
    of include/allwave_cuda.h with the same arity and the same pointer / integer widths, and the same struct layouts"""
    hdr = open(os.path.join(ROOT, "include", "allwave_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    rs = open(os.path.join(ROOT, "bindings", "allwave-cuda-sys", "src", "lib.rs")).read()
    rs = re.sub(r"//.*", "", rs)
    cdecl = {}
    for ret, name, args in re.findall(r"^\s*([A-Za-z_][\w \*]*?)\s*\b(aw_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.M | re.S):
        if "typedef" in ret:
            continue
        params = []
        for a in [x.strip() for x in args.replace("\n", " ").split(",")]:
            if a in ("", "void"):
                continue
            is_array = bool(re.search(r"\[\d*\]$", a))
            a = re.sub(r"\[\d*\]$", "", a)
            m = re.match(r"(.*?)(\b[A-Za-z_]\w*)$", a)
            ty = m.group(1).strip() if m and m.group(1).strip() else a
            params.append(ty + "*" if is_array else ty)
        cdecl[name] = (ret.strip(), params)
    ext = re.search(r'extern "C" \{(.*?)\n\}', rs, flags=re.S).group(1)
    rdecl = {}
    for name, args, ret in re.findall(r"pub fn (aw_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*([^;]+?))?\s*;", ext, flags=re.S):
        params = [re.sub(r"\s+", " ", a.split(":", 1)[1].strip()) for a in args.replace("\n", " ").split(",") if ":" in a]
        rdecl[name] = ((ret or "()").strip(), params)
    assert set(cdecl) == set(rdecl) == set(aw._cabi.EXPORTS), (set(cdecl) ^ set(rdecl), set(cdecl) ^ set(aw._cabi.EXPORTS))
    for name, (cret, cparams) in cdecl.items():
        rret, rparams = rdecl[name]
        want = [_c_type_to_rust(p) for p in cparams]  # arrays in prototypes (uint64_t out[8]) decay to pointers
        assert rparams == want, (name, rparams, want)
        assert rret == _c_type_to_rust(cret), (name, rret, cret)
    # struct fields, in order
    for cname in ("aw_params", "aw_pair", "aw_result"):
        cbody = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), hdr, flags=re.S).group(1)
        cfields = []
        for decl in [d.strip() for d in cbody.split(";") if d.strip()]:
            m = re.match(r"(.*?)(\b\w+(?:\s*,\s*\w+)*)$", decl)
            for f in [x.strip() for x in m.group(2).split(",")]:
                cfields.append((f, _c_type_to_rust(m.group(1))))
        rbody = re.search(r"pub struct %s \{(.*?)\}" % cname, rs, flags=re.S).group(1)
        rfields = [(n, re.sub(r"\s+", " ", t.strip())) for n, t in re.findall(r"pub (\w+)\s*:\s*([^,]+),", rbody)]
        assert rfields == cfields, (cname, rfields, cfields)
    # status / flag constants
    for cname, val in re.findall(r"\b(AW_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr) + re.findall(r"#define (AW_FLAG_[A-Z_]+) (\d+)u", hdr):
        m = re.search(r"pub const %s: \w+ = (-?\d+);" % cname, rs)
        assert m and int(m.group(1)) == int(val), cname


def test_validate_cigar_alignment_host():
    """wfa::validate_cigar_alignment of the C++ mirror: the cases and messages of src/wfa.rs:105-176 (WFA2 letters: I consumes
    the reference, D the query; '=' is accepted beside M and X)"""
    from allwave_b200 import hostlib as H

    assert H.validate_cigar_alignment(b"MMMXMM", 6, 6) is None
    assert H.validate_cigar_alignment(b"MM=XIIDM", 6, 7) is None          # query: MM=X + D + M = 6, reference: MM=X + II + M = 7
    assert H.validate_cigar_alignment(b"", 0, 0) is None
    assert H.validate_cigar_alignment(b"MMM", 2, 3) == "CIGAR extends beyond sequences at M/=/X op: q_pos=2, r_pos=2, query_len=2, ref_len=3"
    assert H.validate_cigar_alignment(b"MI", 1, 1) == "CIGAR extends beyond reference at I op: r_pos=1, ref_len=1"
    assert H.validate_cigar_alignment(b"MD", 1, 1) == "CIGAR extends beyond query at D op: q_pos=1, query_len=1"
    assert H.validate_cigar_alignment(b"MSM", 3, 3) == "Invalid CIGAR operation: S (0x53)"
    assert H.validate_cigar_alignment(b"MM", 3, 2) == "CIGAR doesn't cover full query: 2 vs 3"
    assert H.validate_cigar_alignment(b"MM", 2, 3) == "CIGAR doesn't cover full reference: 2 vs 3"
