"""dev probe: where the end-to-end (host buffers) time of one C2 batch goes."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import allwave_b200 as aw
from allwave_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
c, ids, seqs, rc = synth.config("C2", n=1000)
n = len(seqs)
pairs = [(i, (i * 7 + 1 + k) % n) for k in range(3) for i in range(n)]
pairs = [(q, t) for q, t in pairs if q != t][:B]
L = aw._cabi.lib()
params = aw.make_params(0, 5, 8, 2, 24, 1)
sa = (C.c_char_p * n)(*seqs); la = (C.c_uint64 * n)(*[len(s) for s in seqs]); ia = (C.c_char_p * n)(*[i.encode() for i in ids])
arr = aw._cabi.make_pairs(pairs)
acc = [0]
def _cb(rp, _u):
    acc[0] += rp.contents.paf_len
    return 0
cb = aw._cabi.RESULT_CB(_cb)
ctx = aw.Context(0)
for it in range(3):
    t0 = time.perf_counter()
    aw._cabi.check(L.aw_load_sequences(ctx._h, n, sa, la, ia), "load")
    t1 = time.perf_counter()
    b = aw.Batch(ctx, params, pairs, orientation=aw.AW_ORIENT_MASH, flags=0)
    t2 = time.perf_counter()
    b.launch()
    import torch; torch.cuda.synchronize()
    t3 = time.perf_counter()
    b.fetch(collect=False)
    t4 = time.perf_counter()
    b.close()
    t5 = time.perf_counter()
    aw._cabi.check(L.aw_align_pairs(ctx._h, C.byref(params), aw.AW_ORIENT_MASH, arr, len(pairs), 0, cb, None), "align")
    t6 = time.perf_counter()
    print(f"it{it}: load {1e3*(t1-t0):.1f} ms | create {1e3*(t2-t1):.1f} | launch+sync {1e3*(t3-t2):.1f} (kernel {b.kernel_ms() if False else 0}) | fetch {1e3*(t4-t3):.1f} | destroy {1e3*(t5-t4):.1f} | align_pairs(cb) {1e3*(t6-t5):.1f}")
