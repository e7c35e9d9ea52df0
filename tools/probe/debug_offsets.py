import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
lg = importlib.import_module("l-giremi_b200")
synth = importlib.import_module("l-giremi_b200.synth")
enc = importlib.import_module("l-giremi_b200.encode")
rng = np.random.default_rng(3)
def unit(S, R, cov=0.6):
    a, k = synth.draw_alleles(rng, 1, S, R, cov)
    return enc.EncodedUnit(list(range(S)), [("mismatch", "snp", "het_snp")[int(x)] for x in k[0]], synth.labels_from_alleles(a[0]))
ctx = lg.Context(0)
eus = [unit(S, R) for S, R in [(2, 6), (7, 33), (50, 200), (60, 256), (64, 256), (65, 40), (70, 300), (33, 1000), (130, 17), (20, 2100)]]
eus.append(enc.EncodedUnit([], [], np.zeros((0, 0), np.uint8)))
eus.append(enc.EncodedUnit([5], ['het_snp'], np.full((1, 9), 2, np.uint8)))
lab = rng.choice(np.array([0, 1, 2, 255], np.uint8), size=(12, 150), p=[0.3, 0.3, 0.3, 0.1])
eus.append(enc.EncodedUnit(list(range(12)), ['het_snp', 'mismatch'] * 6, lab))
pb = lg.pack_units(eus)
mode = lg.MODE_ALL_PAIRS | lg.MODE_EMIT_COUNTS
ctx.set_dense_threshold(2, 1)
d = lg.mi_step_batched(pb, 6, mode, ctx=ctx, n_chunks=1)
print("single  off", d.unit_rec_off.tolist())
for chunks in (1, 2, 3):
    p = lg.Pipeline(ctx, pb, chunks)
    for extra, kw in ((0, {}), (lg.MODE_SPLIT_RECORDS, dict(packed=True))):
        r = p.step(6, mode | extra, **kw)
        print("pipe", chunks, "split" if extra else "plain", "off", r.unit_rec_off.tolist(), "nrec", r.n_records)
    p.close()
