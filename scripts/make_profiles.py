"""Turns the outputs of scripts/final_run.sh (gpurun_out/final/) into the tracked summaries under profiles/."""
import collections, csv, io, json, shutil, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
src = ROOT / "gpurun_out" / "final"
dst = ROOT / "profiles"


def bench_line(w):
    return [l for l in open(src / f"bench_{w}.log") if l.startswith("{")][-1]


lines = [bench_line(w) for w in ("c1", "c2", "c3", "c4", "c5", "ref_c4")]
(dst / "r01_final_bench_lines.jsonl").write_text("".join(lines))
c4 = json.loads(bench_line("c4"))
shutil.copy(src / "launches_c4.csv", dst / "r01_e_launches_c4_cluster_solver.csv")
rows = [r for r in csv.reader(open(src / "launches_c4.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", ""); v = float(r[-1].replace(",", "")); unit = r[-2]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for k, a in agg.items() if "probe" not in k)
tab = "\n".join(f"| {k} | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {100 * a[1] / tot:.1f}% |"
                for k, a in sorted(agg.items(), key=lambda x: -x[1][1]) if "probe" not in k)
phase = [l for l in open(src / "phase.log") if l.startswith("c4")][-1].strip()
(dst / "r01_e_launches_c4_cluster_solver.md").write_text(f"""# Round 1, final state — ncu launch list, `python bench.py --steps 2 --warmup 3 --workload c4 --no-cpu-baseline`

`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (first 400 launches; per-launch times are cold-cache and
serialised: compare SHARES with bench.py's CUDA-event numbers, not absolutes).  Raw list: `r01_e_launches_c4_cluster_solver.csv`.
`k_dfma_probe` (2 launches, the fp64-peak probe `bench.py` runs OUTSIDE the timed region) is left out of the shares.

| kernel | launches | total us | avg us | share |
|---|---:|---:|---:|---:|
{tab}

bench.py (same box, CUDA events, graph replay, L2 flushed between iterations): {c4['ms_per_step']:.3f} ms per LM iteration, of which the
linearise+Schur pass is {c4['roofline']['kernel_ms']:.3f} ms; per-phase CUDA-event timing (`scripts/phase_times.py`, serialised, ms):
`{phase}`.
The shares agree: the lineariser is the largest phase, then the band solve, then the back-substitution.

History of the c4 iteration in this round (same measurement): 2.92 ms lineariser alone with global atomics -> 0.773 ms per iteration
(tiled lineariser + block band solver, `r01_b`) -> 0.645 ms (panel-warp lookahead, `r01_d`) -> 0.435 ms (two-CTA cluster solver,
speculative 6x6 factor, uniform row solves, separator by continued elimination) -> 0.39 ms (ring refill without index arithmetic,
cp.async prefetch in the lineariser, epilogue over camera slices, batched staging loads).
""")


def raw(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return {h: (v, u) for h, u, v in zip(r[0], r[1], r[2])}


want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = []
reports = {"lin": raw(src / "lin_c4.ncu-rep"), "band": raw(src / "band_c4.ncu-rep")}
for name, title in (("lin", "k_lin_tile2<4,128,4,1> — linearise + Schur pass (c4)"),
                    ("band", "k_chol_banded_c2<2> — two-CTA cluster band Cholesky (c4, n = 1188, half-bandwidth 29)")):
    d = reports[name]
    out.append(f"## {title}\n\n| metric | value | unit |\n|---|---:|---|")
    out += [f"| `{w}` | {d[w][0]} | {d[w][1]} |" for w in want if w in d]
    st = {k: v for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
    out.append("\nWarp stall reasons (cycles per issued instruction):\n")
    for k, v in sorted(st.items(), key=lambda kv: -float(kv[1][0] or 0))[:9]:
        out.append(f"* {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {float(v[0]):.2f}")
    out.append("")
lin = reports["lin"]


def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


rd = to_bytes(*lin["dram__bytes_read.sum"]); wr = to_bytes(*lin["dram__bytes_write.sum"])
(dst / "r01_f_final_ncu_full_lin_and_band.md").write_text(f"""# Round 1 (final) — `ncu --set full --clock-control none --import-source on` of the two dominant kernels

Workload c4 (200 keyframes, 200k points, 999 062 observations), one B200.  Commands (each run to exit 0 first without ncu):
`python scripts/lin_times.py c4` under `ncu ... -k regex:k_lin_tile2 -s 5 -c 1` and `python scripts/phase_times.py c4` under
`ncu ... -k regex:k_chol_banded_c2 -s 3 -c 1` (`scripts/final_run.sh`; this file is written by `scripts/make_profiles.py`).
Durations under ncu are not bench values.

""" + "\n".join(out) + f"""
## Reading

* Lineariser: DRAM traffic per launch {rd / 1e6:.1f} MB (read) + {wr / 1e6:.1f} MB (write), below the algorithmic 59.5 MB of SURVEY.md §8(d); DRAM
  and L2 are a few % of peak, the fp64 pipe is ~20 % active with 2 warps per scheduler (255 registers per thread): the kernel is bound
  by issue/latency of its fp64 dependency chains, not by memory.  Source-level view (`--page source`): phase 1 (per-observation
  linearisation, 1 150 warp instructions per chunk, 480 of them fp64) holds 66 % of the stall samples, the DMMA phase 19 % (it runs at
  the fp64 pipe's rate: one DMMA.8x8x4 occupies it for 16 cycles), flushes and setup the rest.  With the register prefetch the single
  hottest lines were spill stores of just-loaded values (14 % of the samples); the cp.async staging removed them (long-scoreboard
  stalls 1.5 -> 0.5 cycles per issue).  Three CTAs per SM by a 168-register cap: 0.28 ms instead of 0.19 (spills).  `bench.py` copies
  the traffic figure from `profiles/ncu_traffic.json`.
* Band solver: one cluster of two CTAs on two SMs, everything in shared memory (DRAM traffic is the factor rows out and back);
  a latency chain of 96 + 5 block steps forward and 102 blocks backward per CTA (`scripts/band_timing.py` gives the phase split).
""")
json.dump({"c4": {"dram_bytes_per_launch": int(round(rd + wr)), "kernel": "k_lin_tile2<4,128,4,1>",
                  "source": "profiles/r01_f_final_ncu_full_lin_and_band.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}},
          open(dst / "ncu_traffic.json", "w"), indent=1)
print("profiles written; c4 traffic", int(round(rd + wr)))
