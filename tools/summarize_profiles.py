"""Turn the ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/."""
import collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
ROUND = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs(OUT, exist_ok=True)

# ---- 1. launch list: per-kernel device time (cold cache, serialised: compare SHARES)
path = os.path.join(GO, f"launches_{ROUND}.csv")
if os.path.exists(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    rows = list(csv.DictReader(lines))
    # keep only one inference step: the launches between the first and second embed_kernel
    idx = [i for i, r in enumerate(rows) if "embed_kernel" in r["Kernel Name"]]
    # dual-chain inference: the batch is split in two halves that run as two launch chains -> two embed kernels per step
    step = rows[idx[-3]:idx[-1]] if len(idx) >= 3 else rows
    single_chain = len(idx) >= 2 and (idx[-1] - idx[-2]) <= 100   # round 2: one launch chain of ~75 kernels per step
    if single_chain:
        step = rows                                              # the whole -s / -c window (consecutive steps)
    for r in step:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("vb::", "")
        key = (name, r["Grid Size"], r["Block Size"])
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, f"launches_{ROUND}.md"), "w") as f:
        if single_chain:
            per_step = idx[-1] - idx[-2]
            f.write(f"# ncu launch list, VAENAR.inference at C2 (B16, T_text 148, T_mel 870), one launch chain -- {ROUND}\n\n"
                    "`ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 160 python bench.py --steps 2 --warmup 1 "
                    "--skip-cpu --no-train --no-audio --inflight 1`\n"
                    f"(the {len(step)} launches after the warm-up = consecutive steps of {per_step} launches incl. the noise kernel; "
                    "per-launch times are cold-cache and serialised: compare SHARES).\n\n"
                    f"{len(step)} launches, {tot:.0f} us summed device time = {tot * per_step / len(step):.0f} us per step serialised.\n\n"
                    "| kernel | grid | block | launches | total us | avg us | share |\n|---|---|---|---|---|---|---|\n")
        else:
          f.write(f"# ncu launch list, one VAENAR.inference step at C2 (B16, T_text 148, T_mel 870) -- {ROUND}\n\n")
        if not single_chain:
          f.write("`ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 1`; "
                "per-launch times are cold-cache and serialised, so compare SHARES, not absolutes.\n\n")
          f.write(f"{len(step)} launches (both launch chains of the step: batch halves 8 + 8 run concurrently on two streams; ncu "
                f"serialises them), {tot:.0f} us summed device time.\n\n| kernel | grid | block | launches | total us | avg us | share |\n|---|---|---|---|---|---|---|\n")
        for (name, grid, blk), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {name} | {grid} | {blk} | {n} | {t:.1f} | {t / n:.1f} | {t / tot:.3f} |\n")
        by = collections.defaultdict(float)
        for (name, _, _), (n, t) in agg.items():
            by[name] += t
        f.write("\n| kernel (all grids) | total us | share |\n|---|---|---|\n")
        for k, t in sorted(by.items(), key=lambda kv: -kv[1]):
            f.write(f"| {k} | {t:.1f} | {t / tot:.3f} |\n")
    print("wrote launches summary:", len(step), "launches", f"{tot:.0f} us")

# ---- 1b. launch list of one train_step (tools/ncu_train.py under ncu --profile-from-start off)
path = os.path.join(GO, f"launches_train_{ROUND}.csv")
if os.path.exists(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("vb::", "")
        if "at::" in name:
            name = "torch:" + re.sub(r"<.*", "", name).split("::")[-1]
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, f"launches_train_{ROUND}.md"), "w") as f:
        f.write(f"# ncu launch list, one train_step at C3 (B32, T_text 148, T_mel 870, rf 2) -- {ROUND}\n\n"
                "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python tools/ncu_train.py 32 2` "
                "(forward with tape + backward + Adam + operand re-pack; the weight-gradient stream is serialised by ncu). "
                "Per-launch times are cold-cache and serialised: compare SHARES.\n\n")
        f.write(f"{len(rows)} launches, {tot:.0f} us summed device time.\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {name} | {n} | {t:.1f} | {t / n:.1f} | {t / tot:.3f} |\n")
    print("wrote train launch summary:", len(rows), "launches", f"{tot:.0f} us")

# ---- 2. full-set captures -> key metrics per kernel
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


traffic = {}
KREG = {"attn": "attention_tc", "gemm": "gemm_tc", "gemm2": "gemm_tc", "wgrad": "wgrad_tc", "attn_bwd": "attn_bwd_d"}
for tag, labels in (("attn", ["self-attention causal (B16,H4,Tq=Tk=435)"] * 2 + ["cross-attention (Tq 435, Tk 148)"] * 2 +
                      ["decoder cross-attention with alignments output (Tq 435, Tk 148)"]),
                    ("gemm", ["FFN dense1 + bias + relu (M6960,K256,N1024), BLOCK_N 128"] * 2 +
                     ["FFN dense2 + bias + residual + LayerNorm (M6960,K1024,N256), BLOCK_N 256 single CTA"] * 2),
                    ("gemm2", ["two CTAs per SM: dgrad through ffn.dense1, accumulated (M13920,K1024,N256), F_RES|F_OUT_F32"] * 2 +
                     ["two CTAs per SM: (M13920,K256,N1024), fp32 out"] * 2),
                    ("wgrad", ["FFN dense1 weight gradient dW[256,1024] over 13920 tokens (C3)"] * 2 +
                     ["att_proj weight gradient dW[512,256] ([x ; ctx] concat) over 13920 tokens"] * 2),
                    ("attn_bwd", ["causal self-attention backward (B32,H4,T435): dK/dV kernel", "same: dQ kernel"] * 2 +
                     ["cross-attention backward (Tq 435, Tk 148): dK/dV kernel", "same: dQ kernel"] * 2)):
    rep = os.path.join(GO, f"prof_{tag}_{ROUND}.ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, rows = raw(rep)
    with open(os.path.join(OUT, f"{tag}_{ROUND}.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none: {tag} kernels at C2 shapes -- {ROUND}\n\n"
                f"`ncu --set full --clock-control none --import-source on -k regex:{KREG[tag]} "
                f"python tools/prof_kernels.py {tag}` (block-level C-ABI hooks, same kernels/shapes as the model).\n")
        for k, r in enumerate(rows):
            name = r[hdr.index("Kernel Name")]
            f.write(f"\n## launch {k}: {labels[k] if k < len(labels) else ''}\n`{name[:110]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for w in WANT:
                if w in hdr:
                    f.write(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |\n")
            dr = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            dw = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            f.write(f"| dram traffic (read+write) | {dr + dw:.0f} | byte |\n")
            if tag == "gemm" and k == 1:
                traffic["gemm_plain"] = dr + dw
            if tag == "gemm" and k == 3:
                traffic["gemm_ln"] = dr + dw
            if tag == "attn" and k == 1:
                traffic["attn_self"] = dr + dw
            if tag == "attn" and k == 3:
                traffic["attn_cross"] = dr + dw
            if tag == "attn" and k == 4:
                traffic["attn_cross_ali"] = dr + dw
            if tag == "wgrad" and k == 1:
                traffic["wgrad"] = dr + dw
            if tag == "attn_bwd" and k == 2:
                traffic["attn_bwd_dkdv_self"] = dr + dw
            if tag == "attn_bwd" and k == 3:
                traffic["attn_bwd_dq_self"] = dr + dw
    print("wrote", tag)
if traffic:
    old = {}
    try:
        old = json.load(open(os.path.join(OUT, "traffic.json")))
    except Exception:
        pass
    old.update(traffic)
    traffic = old
    json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    print(traffic)
