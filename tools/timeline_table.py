"""Per-warp table of the phase stamps of the instrumented sweeps (GSCAN_TIMELINE=<prefix> dumps <prefix>.fwd.bin /
<prefix>.bwd.bin: [T][32 stamps][16 warps] clock64 values of CTA 0, lane 0 of every warp).  Rows = stamps, columns =
warps; entries = average cycles after the step's first stamp.  Usage: timeline_table.py file.bin"""
import sys
import numpy as np
NS = 32
h = np.fromfile(sys.argv[1], dtype=np.int64)
T = h.size // (NS * 16)
h = h.reshape(T, NS, 16).astype(np.float64)
h[h == 0] = np.nan
rev = h[2, 0, 0] > h[T - 3, 0, 0]
base = np.nanmin(h[:, 0, :], axis=1)
rel = np.nanmean((h - base[:, None, None])[5:T - 5], axis=0)
step = np.nanmean(np.abs(np.diff(base[5:T - 5])))
np.set_printoptions(linewidth=250, suppress=True)
print(f"{'backward' if rev else 'forward'} sweep, {T} steps, {step:.0f} cycles per step; rows = stamps, columns = warps 0..15")
for k in range(NS):
    if np.all(np.isnan(rel[k])):
        continue
    print(f"{k:2d} " + " ".join("    ." if np.isnan(v) else f"{v:5.0f}" for v in rel[k]))
