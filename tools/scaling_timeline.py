"""profiles/r02_scaling_timeline.md from the bench lines profiles/r02_bench_n{1,2,4,8}.json."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {n: json.loads(open(os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")).read().strip().splitlines()[-1]) for n in (1, 2, 4, 8)}
d1, d8 = rows[1], rows[8]
out = ["# Round 2 — strong scaling of the C5 solve at 1 / 2 / 4 / 8 B200, with the per-N timeline\n",
       "`bench.py --gpus N --steps 5 --warmup 3` (torchrun, one rank per GPU, peer-memory halo transport), builder-run on the pod's\n"
       "8-GPU box; lines in `profiles/r02_bench_n{1,2,4,8}.json`. 100 BiCGStab iterations per step; `step_share` = CUDA-event time per\n"
       "kernel family inside one extra step (the headline takes the max over ranks, the shares are rank 0's). The boxes differ in\n"
       "their power state (`clocks` in the lines: 1965 MHz uncapped at N = 1 and 8, ≈1925 MHz under `sw_power_cap` at N = 2 and 4).\n",
       "| N | it/s (device) | speed-up | efficiency | it/s e2e (host slices) | e2e speed-up | SpMV ms / launch | SpMV frac of peak | SpMV ms/step | vector ms/step | scalar + all-reduce ms/step | host copies ms/step | SM MHz | parity pre-check (96³ golden) |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for n in (1, 2, 4, 8):
    d = rows[n]
    r = d["roofline"]
    ss = r["step_share"]
    pc = d["parity_check"]
    out.append(f"| {n} | {d['value']:.1f} | {d['value'] / d1['value']:.2f} | {d['value'] / d1['value'] / n:.3f} | {d['e2e']['value']:.1f} | "
               f"{d['e2e']['value'] / d1['e2e']['value']:.2f} | {r['avg_launch_ms']:.3f} | {r['frac']:.3f} | {ss['spmv_ms']:.1f} | {ss['vector_ms']:.1f} | "
               f"{ss['scalar_ms']:.1f} | {d['e2e']['ms_per_step'] - d['ms_per_step']:.1f} | {d['clocks']['sm_mhz']:.0f} | "
               f"{pc['iterations']} its, history bit-identical: {pc['history_bit_identical']} ({pc['ranks']} ranks) |")
s1, s8 = d1["roofline"]["step_share"], d8["roofline"]["step_share"]
ideal = d1["ms_per_step"] / 8
copy8 = d8["e2e"]["ms_per_step"] - d8["ms_per_step"]
copy1 = d1["e2e"]["ms_per_step"] - d1["ms_per_step"]
gb = (d1["e2e"]["h2d_bytes_per_step"] + d1["e2e"]["d2h_bytes_per_step"]) / 1e9
out += ["",
        f"Where the {d8['ms_per_step'] - ideal:.1f} ms per step above ideal (1/8 of the 1-GPU step = {ideal:.1f} ms) go at N = 8:\n",
        f"* **SpMV** {s8['spmv_ms']:.1f} ms vs {s1['spmv_ms'] / 8:.1f} ideal (+{s8['spmv_ms'] - s1['spmv_ms'] / 8:.1f}): the kernel's fixed cost per launch is 12 µs (`tools/spmv_fixed_cost.py`: "
        f"t = 12.3 µs + 9.87 µs per 512² plane, single GPU, no halo) of a {d8['roofline']['avg_launch_ms'] * 1e3:.0f} µs launch; the rest is the partitioned variant — boundary tiles gather "
        f"from the peer-written halo window and select the base pointer per gather (interior tiles skip the select since this round), the producer lanes check the neighbours' flags "
        f"before the first boundary tile — per-byte efficiency {d1['roofline']['frac']:.3f} → {d8['roofline']['frac']:.3f}.",
        f"* **vector kernels** {s8['vector_ms']:.1f} vs {s1['vector_ms'] / 8:.1f} ideal (+{s8['vector_ms'] - s1['vector_ms'] / 8:.1f}): launch ramp of three kernels per iteration on 1/8 of the rows.",
        f"* **reduction points** {s8['scalar_ms']:.1f} vs {s1['scalar_ms']:.1f} ms at N = 1 (+{s8['scalar_ms'] - s1['scalar_ms']:.1f}): three all-reduces per iteration, each = one small kernel that folds the block "
        f"partials, stores the four (hi, lo) pairs into the seven peers' windows and waits for their flags — ≈15 µs more than the single-GPU finish; it absorbs the arrival skew of "
        f"the eight ranks, which the events attribute to this family. The halo put (pack kernel, ≈5 µs per product) is the remaining {d8['ms_per_step'] - s8['spmv_ms'] - s8['vector_ms'] - s8['scalar_ms']:.1f} ms.",
        f"* **e2e arm**: the box is one socket / one NUMA node with 32 vCPUs (`profiles/r02_topo_n8.txt`), so the host copies are bound by what the host memory system feeds eight PCIe links "
        f"at once: {gb:.1f} GB per step in {copy8:.0f} ms = {gb / (copy8 / 1e3):.0f} GB/s aggregate ({copy1:.1f} ms for the same bytes on one link at N = 1 = {gb / (copy1 / 1e3):.0f} GB/s); NUMA binding has "
        f"nothing to bind to here. The device-resident speed-up is {d8['value'] / d1['value']:.2f}×, the end-to-end one {d8['e2e']['value'] / d1['e2e']['value']:.2f}×.",
        "\nThe full 512³ solve takes 640 iterations with the same final residual at every N (`full_solve` in the JSON lines) and the partitioned 96³ case reproduces the committed "
        "exact-dot golden history bit for bit at 1, 2, 4 and 8 ranks."]
open(os.path.join(ROOT, "profiles", "r02_scaling_timeline.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[5:11]))
