"""Turn the ncu reports / launch list that tools/run_prof.sh left in gpurun_out/ into the tracked summaries under
profiles/ (names per round: ROUND=r2 by default).  usage: [ROUND=r2] python tools/make_profiles.py"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PRO = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = os.environ.get("ROUND", "r2")

def summary(rep, dst):
    p = os.path.join(OUT, rep)
    if not os.path.exists(p):
        return None
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), p], capture_output=True, text=True).stdout
    open(os.path.join(PRO, dst), "w").write(txt)
    for line in txt.splitlines():
        if "dram traffic" in line:
            v, unit = line.split(":")[1].split()[:2]
            return float(v) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(unit, 1)
    return None

def sass_hist(rep, dst, div):
    p = os.path.join(OUT, rep)
    if not os.path.exists(p):
        return
    src = subprocess.run(["ncu", "-i", p, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    tmp = os.path.join(OUT, "_src.csv")
    open(tmp, "w").write(src)
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_hist.py"), tmp, str(div)], capture_output=True, text=True).stdout
    open(os.path.join(PRO, dst), "w").write("\n".join(txt.splitlines()[:32]) + "\n")

def launches():
    p = os.path.join(OUT, "launches.csv")
    if not os.path.exists(p):
        return
    rows = list(csv.reader(l for l in open(p) if not l.startswith("==")))
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[mv].replace(",", ""))
        except Exception:
            continue
        tot[r[kn]] += v / (1e3 if r[mu].startswith("n") else 1.0)
        cnt[r[kn]] += 1
    s = sum(tot.values())
    lines = ["launch list of `python bench.py --steps 5 --warmup 3 --no-cpu` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)",
             "kernel | launches | total us | share"]
    for k, v in tot.most_common():
        lines.append("%-70s %5d %12.1f %5.1f%%" % (k[:70], cnt[k], v, 100 * v / s))
    open(os.path.join(PRO, R + "_launches_summary.txt"), "w").write("\n".join(lines) + "\n")
    import shutil
    shutil.copy(p, os.path.join(PRO, R + "_launches_bench.csv"))

launches()
tj = os.path.join(PRO, R + "_traffic.json")
d = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures in this directory "
                 "(%s_*_full.txt); bench.py copies the clFFT figure into roofline.traffic when its launch shape matches" % R}
XE = 32 * 1024 * 1024 * 2 + 1024 * 528 * 8
for rep, dst, key, alg in (("fft.ncu-rep", "fft8192_full.txt", "k_fft_8192pt_x8192vec", 1073741824),
                           ("fft4096.ncu-rep", "fft4096_full.txt", "k_fft_4096pt_x16384vec", 1073741824),
                           ("fftfilt.ncu-rep", "fftfilt_full.txt", "k_fftfilt_256tap_64Mi", 1073741824),
                           ("fir.ncu-rep", "fir_full.txt", "k_fir_256tap_64Mi", 1073741824),
                           ("pfb.ncu-rep", "pfb_full.txt", "k_pfb_64ch_64Mi", 1073741824),
                           ("map1.ncu-rep", "mathconst_full.txt", "k_map1_multiplyconst_64Mi", 1073741824),
                           ("fftcolA.ncu-rep", "fft65536_passA_full.txt", "k_fft_col_passA_65536pt_x1024vec", 1073741824),
                           ("fftcolB.ncu-rep", "fft65536_passB_full.txt", "k_fft_col_passB_65536pt_x1024vec", 1073741824),
                           ("xe_tma.ncu-rep", "xengine_full.txt", "k_xengine_tma_32st_1024ch_1024t", XE),
                           ("xe_batch.ncu-rep", "xengine_batch16_full.txt", "k_xengine_tma_batch16_32st_1024ch_1024t", 16 * XE),
                           ("xe_c32.ncu-rep", "xengine_c32_full.txt", "k_xengine_c32_32st_256ch_1024t", 32 * 256 * 1024 * 8 + 256 * 528 * 8),
                           ("xe_pk.ncu-rep", "xengine_packed_full.txt", "k_xengine_tma_packed_16st_2pol_1024ch_1024t", 16 * 2 * 1024 * 1024 + 1024 * 136 * 4 * 8)):
    b = summary(rep, R + "_" + dst)
    if b:
        d[key] = {"bytes": int(b), "algorithmic_bytes": alg, "capture": "profiles/" + R + "_" + dst}
json.dump(d, open(tj, "w"), indent=2)
sass_hist("fft.ncu-rep", R + "_fft8192_sass_hist.txt", 8192)
sass_hist("xe_batch.ncu-rep", R + "_xengine_batch16_sass_hist.txt", 16)
sass_hist("fftfilt.ncu-rep", R + "_fftfilt_sass_hist.txt", 1)
print(open(os.path.join(PRO, R + "_launches_summary.txt")).read())
