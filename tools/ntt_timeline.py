"""Per-CTA phase timeline of the NTT pass kernel (needs the instrumented build:
   P3GPU_OUT=$PWD/build/libp3gpu_prof.so P3GPU_OBJ=$PWD/build/obj_prof plonky3_b200/csrc/build.sh -DP3GPU_NTT_PROFILE
   P3GPU_LIB=$PWD/build/libp3gpu_prof.so python tools/ntt_timeline.py [w]).
Prints, per launch of one 2^20 x w LDE, the mean duration (us) of: wait for the previous tile's readers, cp.async issue,
load wait, step 1, step 2 (+stores), and the co-residency of CTAs on an SM."""
import os, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import numpy as np

w = int(sys.argv[1]) if len(sys.argv) > 1 else 100
buf = torch.zeros(8 * (1 << 17), dtype=torch.int64, device="cuda")
os.environ["P3GPU_NTT_PROFBUF"] = str(buf.data_ptr())
from plonky3_b200.field import KoalaBear as KB
from plonky3_b200.gpu import default_gpu
gpu = default_gpu(0)
x = torch.randint(0, KB.P, (1 << 20, w), device="cuda", dtype=torch.int32)
for _ in range(2):   # 2 LDEs = 8 launches = the 8 windows; the second LDE (warm) overwrites windows 4..7
    y = gpu.coset_lde_batch(KB.id, x, 1, KB.generator)
torch.cuda.synchronize()
raw = buf.cpu().numpy().reshape(8, -1)
names = ["inverse 0-9", "inverse 10-19", "forward 0-9", "forward 10-19"]
if os.environ.get("P3GPU_NTT_PIPE", "1") != "0":
    # pipelined kernel: slots per (CTA, group, tile): smid/unused, t_start, t_full, t_step1, t_step2
    for li in range(4, 8):
        d = raw[li].reshape(-1, 16, 8)
        rows = d[d[:, :, 4] != 0].astype(np.float64)
        print(f"launch {names[li - 4]}: {len(rows)} stamped group-tiles")
        print(f"   wait for tile {np.mean(rows[:, 2] - rows[:, 1]) / 1e3:6.2f} us (p50 {np.percentile(rows[:, 2] - rows[:, 1], 50) / 1e3:.2f}, p90 "
              f"{np.percentile(rows[:, 2] - rows[:, 1], 90) / 1e3:.2f}) | step1 {np.mean(rows[:, 3] - rows[:, 2]) / 1e3:6.2f} | "
              f"step2+stores {np.mean(rows[:, 4] - rows[:, 3]) / 1e3:6.2f} | total {np.mean(rows[:, 4] - rows[:, 1]) / 1e3:6.2f}")
    sys.exit(0)
b = raw.reshape(8, -1, 16, 8)
names = ["inverse 0-9", "inverse 10-19", "forward 0-9", "forward 10-19"]
for li in range(4, 8):
    d = b[li]
    valid = d[:, :, 5] != 0
    n_cta = int(valid[:, 0].sum())
    rows = d[valid]
    start, issued, loaded, s1, s2 = (rows[:, i].astype(np.float64) for i in (1, 2, 3, 4, 5))
    print(f"launch {names[li - 4]}: {n_cta} CTAs, {len(rows)} stamped tiles")
    print(f"   issue {np.mean(issued - start) / 1e3:7.2f} us | load wait {np.mean(loaded - issued) / 1e3:7.2f} | step1 {np.mean(s1 - loaded) / 1e3:7.2f} | "
          f"step2+stores {np.mean(s2 - s1) / 1e3:7.2f} | tile total {np.mean(s2 - start) / 1e3:7.2f}")
    # per CTA: gap between consecutive tiles (includes the leading barrier)
    gaps = []
    for c in range(d.shape[0]):
        k = int(valid[c].sum())
        for j in range(1, k):
            gaps.append(float(d[c, j, 1]) - float(d[c, j - 1, 5]))
    if gaps:
        print(f"   gap between tiles {np.mean(gaps) / 1e3:6.2f} us; percentiles of load wait: "
              + ", ".join(f"p{q}={np.percentile(loaded - issued, q) / 1e3:.2f}" for q in (10, 50, 90)))
    t0 = rows[:, 1].min()
    span = (rows[:, 5].max() - t0) / 1e3
    print(f"   stamped span {span:.1f} us")
