"""Turns the round's .ncu-rep captures (gpurun_out/prof_<tag>_<workload>.ncu-rep) into the committed evidence:
profiles/<tag>_ncu_<workload>.md (counter summary + per-line instruction profile) and the two tables bench.py copies
into its JSON line: profiles/executed.json and profiles/traffic.json, keyed "<kernel version>/<scene>_<precision>".

    python tools/r2_profiles.py r02v5 ref teapot gopher tex ref64 cube"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary

SCENE = {"ref": ("reference", "fp32"), "ref64": ("reference", "fp64"), "teapot": ("teapot", "fp32"), "gopher": ("gopher", "fp32"),
         "tex": ("textures", "fp32"), "cube": ("cubemap", "fp32"), "env": ("envmap", "fp32"), "teapot64": ("teapot", "fp64")}
NOTE = {"ref": "reference scene 1280x960, 64 spp", "ref64": "reference scene 1280x960, fp64 mode, 32 spp", "teapot": "teapot 1280x960, 256 spp",
        "gopher": "gopher 1280x960, 256 spp", "tex": "textures scene 1280x960, 32 spp, full-size textures", "cube": "cube-map scene (gopher mesh) 1280x960, 32 spp"}

def num(rec, key):
    try:
        return float(rec[key])
    except Exception:
        return None

tag, names = sys.argv[1], sys.argv[2:]
executed, traffic = {}, {}
for kind, table in (("executed", executed), ("traffic", traffic)):
    path = os.path.join(ROOT, "profiles", f"{kind}.json")
    if os.path.exists(path):
        table.update(json.load(open(path)))
for n in names:
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_{n}.ncu-rep")
    recs, units = ncu_summary.load(rep)
    rec = recs[0]
    md = os.path.join(ROOT, "profiles", f"{tag}_ncu_{n}.md")
    with open(md, "w") as f:
        f.write(f"# kernel {tag}: ncu --set full --clock-control none --import-source on, ptk::trace_kernel, workload: {NOTE.get(n, n)}\n\n"
                f"Captured by tools/r2_ncu.sh under gpurun: the same command exited 0 without ncu first.  Reduced here by tools/r2_profiles.py.\n\n")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, md], stdout=subprocess.DEVNULL, check=True)
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "40"], stdout=subprocess.PIPE, text=True).stdout
    with open(md, "a") as f:
        f.write("\n## Per-source-line warp instructions (top 40)\n```\n" + lines + "```\n")
    scene, prec = SCENE[n]
    key = f"{tag}/{scene}_{prec}"
    executed[key] = {"issue_pct": num(rec, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "fma_pipe_pct": num(rec, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                     "alu_pipe_pct": num(rec, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                     "fp64_pipe_pct": num(rec, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                     "lanes": num(rec, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                     "l1_hit_pct": num(rec, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num(rec, "lts__t_sector_hit_rate.pct"),
                     "local_loads": num(rec, "sass__inst_executed_local_loads"),
                     "source": f"profiles/{tag}_ncu_{n}.md ({NOTE.get(n, n)})"}
    rd, wr = num(rec, "dram__bytes_read.sum"), num(rec, "dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tb = rd * scale.get(units.get("dram__bytes_read.sum"), 1) + wr * scale.get(units.get("dram__bytes_write.sum"), 1)
    traffic[key] = {"bytes_per_launch": tb, "algorithmic_bytes_per_launch": 1280 * 960 * 40,
                    "source": f"profiles/{tag}_ncu_{n}.md: dram__bytes_read.sum + dram__bytes_write.sum, one launch ({NOTE.get(n, n)}; framebuffer traffic is per pixel, independent of spp)"}
    print(key, executed[key]["issue_pct"], executed[key]["lanes"], tb / 1e6, "MB")
json.dump(executed, open(os.path.join(ROOT, "profiles", "executed.json"), "w"), indent=1)
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
