#!/usr/bin/env python3
"""Static issue model of the tile kernel's step loop -- screens source variants on the CPU before GPU time is spent.

    python profiles/tools/sass_issue_model.py [--tag NAME] [--warps 3] [--keep] [-- extra nvcc flags, e.g. -DBR2_TILE_SCALAR_MASKS=0]

What it does (no GPU needed; about 20 s):
  1. compiles ONE instantiation of ``br2_step_tile_kernel`` (profiles/tools/one_kernel.cu; default: the single-segment
     kernel <2, 64, 6>; ``-DK_NT=128 -DK_MB=3 -DK_EX=true -DK_CL=false -DK_SPLIT=false`` selects the double assembly's) to a cubin;
  2. reads ``cuobjdump -sass`` (instruction + control word: stall count, write / read scoreboard, wait mask) and
     ``nvdisasm -gi`` (source line of every instruction, inlining chain) of it;
  3. finds the step loop (largest backward branch spanning 800..4000 instructions), drops the basic blocks of the damper's
     three-exponential path (uniform branch on ``p.tile.round_section``; BR2 rods have round sections) so that the body is
     what a warp executes;
  4. replays the body for W warps that share one scheduler: one issue per cycle, the stall counts ptxas wrote, scoreboard
     waits with nominal latencies (LDS 30, LDC 40, MUFU 25 cycles ...), an fp64 pipe that accepts one warp instruction
     every 2 cycles (B200: 16 lanes per sub-partition), oldest-ready-first arbitration;
  5. prints cycles per step, the fp64 pipe's share, the instruction mix, stall cycles per opcode, the source lines with
     the most non-fp64 instructions and a what-if table (an instruction class removed).

Calibration (single_br2 x 4096, 168-register kernel, 6 CTAs of 2 warps per SM = 3 warps per scheduler; measured =
CUDA-event time per launch x 1.965 GHz / (4.61 work items per CTA x 2000 steps)): see profiles/r02_sass_issue_model.txt.
The model is a SCREEN: it reproduces the measured step time to about 5 % and the sign of the A/B results recorded there;
differences under ~1.5 % are below its resolution and still go to the GPU.
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "br2-simulator_b200", "csrc")
FP64 = ("DFMA", "DMUL", "DADD")
LATENCY = {"LDS": 30, "LDC": 40, "LDCU": 40, "MUFU": 25, "LDG": 600, "LD": 600, "LDL": 40, "STS": 10, "STL": 10, "STG": 10,
           "ST": 10, "ATOM": 600, "ATOMS": 60, "SYNCS": 60, "BAR": 40, "S2R": 30, "I2F": 25, "F2F": 25, "F2I": 25, "DSETP": 12,
           "SHFL": 30, "STAS": 30, "RED": 30, "CCTL": 30, "MEMBAR": 100, "S2UR": 30, "R2UR": 15}


def opcode(text):
    return re.sub(r"^@!?U?P\d+\s+", "", text).split()[0].split(".")[0]


def build(tag, flags, out_dir):
    cubin = os.path.join(out_dir, tag + ".cubin")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-I", os.path.join(ROOT, "include"),
           "-I", CSRC, "-cubin", "-o", cubin, *flags, os.path.join(ROOT, "profiles", "tools", "one_kernel.cu")]
    subprocess.check_call(cmd)
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
    lines = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True, check=True).stdout
    usage = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True, text=True, check=True).stdout
    return sass, lines, " ".join(re.findall(r"(?:REG|STACK|SHARED):\d+", usage))


def parse_sass(sass):
    """[{addr, text, stall, wb, rb, wait}] from ``cuobjdump -sass`` (the control word is the second 64-bit word's bits 41..)."""
    ins, rows = [], sass.split("\n")
    first = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/")
    second = re.compile(r"^\s+/\* (0x[0-9a-f]{16}) \*/")
    for a, b in zip(rows, rows[1:]):
        m, m2 = first.match(a), second.match(b)
        if m and m2:
            c = int(m2.group(1), 16) >> 41
            ins.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=c & 0xF, wb=(c >> 5) & 7, rb=(c >> 8) & 7,
                            wait=(c >> 11) & 0x3F))
    return ins


def step_loop(ins):
    loops = []
    for d in ins:
        m = re.search(r"\bBRA(?:\.U)?\s+(?:\S+,\s*)?(0x[0-9a-f]+)", d["text"])
        if m and int(m.group(1), 16) < d["addr"]:
            loops.append((d["addr"] - int(m.group(1), 16), int(m.group(1), 16), d["addr"]))
    cand = sorted(L for L in loops if 800 <= L[0] // 16 + 1 <= 4000)
    if not cand:
        sys.exit("no step loop found")
    return cand[-1][1], cand[-1][2]


def source_lines(listing, lo, hi):
    """address -> line of tile_step the instruction belongs to (outermost frame inside tile_step), and the basic blocks."""
    src = open(os.path.join(CSRC, "br2_tile.cuh")).read().split("\n")
    t0 = next(i + 1 for i, l in enumerate(src) if "void tile_step(" in l)
    t1 = next(i + 1 for i, l in enumerate(src) if i > t0 and "int kMinBlocks" in l)
    skip = {i + 1 for i, l in enumerate(src) if "M::exp(e)" in l or "(e * e <= BR2_E2_MAX)" in l or "strain * UNI(BR2_FC_LN_CR + i)" in l}
    file_re = re.compile(r'//## File "([^"]+)", line (\d+)')
    ins_re = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);")
    pending, chain, blocks, cur, line_of = [], [], [], [], {}
    for row in listing.split("\n"):
        m = file_re.search(row)
        if m:
            pending.append((m.group(1), int(m.group(2))))
            continue
        m = ins_re.match(row)
        if m:
            if pending:
                chain, pending = pending, []
            addr = int(m.group(1), 16)
            line_of[addr] = next((l for f, l in chain if f.endswith("br2_tile.cuh") and t0 <= l <= t1), 0)
            cur.append((addr, m.group(2).strip()))
            if re.search(r"\bBRA\b", m.group(2)):
                blocks.append(cur)
                cur = []
        elif row.startswith(".L_") and cur:
            blocks.append(cur)
            cur = []
    if cur:
        blocks.append(cur)
    dropped = set()
    for blk in blocks:  # the three-exponential damper path: blocks holding fp64 arithmetic of those source lines
        if any(line_of.get(a) in skip and text.startswith(("DMUL", "DFMA")) for a, text in blk):
            dropped.update(a for a, _ in blk)
    return line_of, dropped, src


def simulate(body, warps, steps=6):
    state = [dict(pc=(w * len(body)) // warps, t=7 * w, sb=[0] * 6, done=0) for w in range(warps)]
    total, pipe_free, now, issued, finish = steps * len(body), 0, 0, 0, [0] * warps
    ops = [opcode(d["text"]) for d in body]
    while any(w["done"] < total for w in state):
        best = None
        for w in state:
            if w["done"] >= total:
                continue
            d, ready = body[w["pc"]], w["t"]
            for b in range(6):
                if d["wait"] >> b & 1:
                    ready = max(ready, w["sb"][b])
            if ops[w["pc"]] in FP64:
                ready = max(ready, pipe_free)
            if best is None or ready < best[0]:
                best = (ready, w)
        ready, w = best
        now = max(now + 1, ready) if issued else ready
        d, op = body[w["pc"]], ops[w["pc"]]
        if op in FP64:
            pipe_free = now + 2
        if d["wb"] != 7:
            w["sb"][d["wb"]] = now + LATENCY.get(op, 25)
        if d["rb"] != 7:
            w["sb"][d["rb"]] = max(w["sb"][d["rb"]], now + 8)
        w["t"] = now + max(d["stall"], 1)
        w["pc"] = (w["pc"] + 1) % len(body)
        w["done"] += 1
        issued += 1
        if w["done"] == total:
            finish[state.index(w)] = now
    return max(finish) / steps


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--tag", default="variant")
    ap.add_argument("--warps", type=int, default=3, help="warps per scheduler (12 resident warps per SM / 4)")
    ap.add_argument("--out", default="/tmp/br2_sass_model")
    ap.add_argument("--brief", action="store_true")
    ap.add_argument("flags", nargs="*", help="extra nvcc flags (after --)")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    sass, listing, usage = build(args.tag, args.flags, args.out)
    ins = parse_sass(sass)
    lo, hi = step_loop(ins)
    line_of, dropped, src = source_lines(listing, lo, hi)
    body = [d for d in ins if lo <= d["addr"] <= hi and d["addr"] not in dropped]
    mix = collections.Counter(opcode(d["text"]) for d in body)
    n_fp64 = sum(mix[k] for k in FP64)
    base = simulate(body, args.warps)
    print("%s  [%s]  flags: %s" % (args.tag, usage, " ".join(args.flags) or "-"))
    print("step loop 0x%x..0x%x: %d instructions per warp and step (%d fp64: %d DFMA %d DMUL %d DADD), stall counts sum to %d cycles"
          % (lo, hi, len(body), n_fp64, mix["DFMA"], mix["DMUL"], mix["DADD"], sum(d["stall"] for d in body)))
    for w in sorted({1, 2, args.warps, 4}):
        c = simulate(body, w)
        print("  %d warp(s) per scheduler: %6.0f cycles per step, fp64 pipe %4.1f %% busy" % (w, c, 100.0 * w * n_fp64 * 2 / c))
    if args.brief:
        return
    print("instruction mix:", ", ".join("%s %d" % kv for kv in mix.most_common(24)))
    stall = collections.Counter()
    for d in body:
        stall[opcode(d["text"])] += d["stall"]
    print("stall cycles by opcode:", ", ".join("%s %d" % kv for kv in stall.most_common(14)))
    print("what-if (instruction class removed from the body, %d warps):" % args.warps)
    classes = [("FSEL / SEL", ("FSEL", "SEL")), ("IMAD / LOP3 / R2UR / UMOV / MOV", ("IMAD", "LOP3", "R2UR", "UMOV", "MOV")),
               ("LDL / STL (spills)", ("LDL", "STL")), ("LDC / LDCU", ("LDC", "LDCU")), ("DSETP", ("DSETP",)),
               ("MUFU / FRND", ("MUFU", "FRND")), ("LDS / STS", ("LDS", "STS"))]
    for name, ops in classes:
        kept = [d for d in body if opcode(d["text"]) not in ops]
        c = simulate(kept, args.warps)
        print("  without %-34s %4d instructions  %6.0f cycles  %+5.1f %%" % (name, len(kept), c, 100.0 * (base / c - 1.0)))
    kept = [d for d in body if opcode(d["text"]) in FP64]
    c = simulate(kept, args.warps)
    print("  %-42s %4d instructions  %6.0f cycles  %+5.1f %%" % ("fp64 arithmetic only", len(kept), c, 100.0 * (base / c - 1.0)))
    per_line = collections.defaultdict(collections.Counter)
    for d in body:
        per_line[line_of.get(d["addr"], 0)][opcode(d["text"])] += 1
    rows = sorted(((sum(v for k, v in c.items() if k not in FP64 + ("LDS", "STS")), line, c) for line, c in per_line.items()), reverse=True)
    print("source lines of tile_step with the most instructions that are neither fp64 arithmetic nor shared-memory accesses:")
    for count, line, c in rows[:16]:
        detail = ", ".join("%s %d" % (k, v) for k, v in c.most_common() if k not in FP64 + ("LDS", "STS"))
        print("  %3d  line %4d  %-60s | %s" % (count, line, detail[:60], src[line - 1].strip()[:80] if line else "(loop control, address arithmetic)"))


if __name__ == "__main__":
    main()
