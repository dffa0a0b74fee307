"""Turn the ncu reports / launch list that tools/run_prof.sh left in gpurun_out/ into the tracked summaries under
profiles/ (round-1 names).  usage: python tools/make_profiles.py"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PRO = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

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
    open(os.path.join(PRO, "r1_launches_summary.txt"), "w").write("\n".join(lines) + "\n")
    import shutil
    shutil.copy(p, os.path.join(PRO, "r1_launches_bench.csv"))

launches()
tj = os.path.join(PRO, "r1_traffic.json")
d = json.load(open(tj))
for rep, dst, key, alg in (("fft.ncu-rep", "r1_fft8192_full.txt", "k_fft_8192pt_x8192vec", 1073741824),
                           ("fftfilt.ncu-rep", "r1_fftfilt_full.txt", "k_fftfilt_256tap_64Mi", 1073741824),
                           ("pfb.ncu-rep", "r1_pfb_full.txt", "k_pfb_64ch_64Mi", 1073741824),
                           ("xe_tma.ncu-rep", "r1_xengine_full.txt", "k_xengine_tma_32st_1024ch_1024t", 71434240)):
    b = summary(rep, dst)
    if b:
        d[key] = {"bytes": int(b), "algorithmic_bytes": alg, "capture": "profiles/" + dst}
json.dump(d, open(tj, "w"), indent=2)
sass_hist("fft.ncu-rep", "r1_fft8192_sass_hist.txt", 8192)
sass_hist("xe_tma.ncu-rep", "r1_xengine_sass_hist.txt", 1)
print(open(os.path.join(PRO, "r1_launches_summary.txt")).read())
