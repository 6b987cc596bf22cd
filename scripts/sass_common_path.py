"""Counts the SASS instructions on the common path of k_align's hot loop (no GPU needed).

The loop is issue-bound, so instructions per candidate on the path every candidate takes is the number to drive down
(profiles/r01_align_ncu_summary.md, row `final2`: 96.5 -> 80.0).  The script compiles csrc/align_kernel.cu to a cubin with the
given extra nvcc flags, finds the ring-stage loop of k_align<false, false> (the backward branch whose body holds the bulk
copy and 32 byte gathers = 8 candidates per lane), and walks the shortest path through one iteration: rare blocks
(warp-uniform branches to slots outside / near the boundary) are longer and therefore not on it.

usage: python scripts/sass_common_path.py [-DVORS_WARPS=12 ...]      -> registers / spills, instructions per candidate,
                                                                        opcode histogram; path written to /tmp/common.sass
"""
import collections, heapq, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "visual-odometry-rs_b200", "csrc", "align_kernel.cu")
CUBIN = "/tmp/vors_align.cubin"


def main():
    flags = [a for a in sys.argv[1:] if a != "--tiled"]
    # which instantiation: the sparse-record kernel k_align<false, false, false> or (--tiled) the tiled dense one <false, false, true>
    KEY = "k_alignILb0ELb0ELb1E" if "--tiled" in sys.argv else "k_alignILb0ELb0ELb0E"
    cmd = ["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
           "--expt-relaxed-constexpr", "-Xptxas", "-v", *flags, "-cubin", SRC, "-o", CUBIN]
    log = subprocess.run(cmd, capture_output=True, text=True, check=True).stderr
    lines = log.splitlines()
    for i, l in enumerate(lines):
        if KEY in l:
            print(" | ".join(x.replace("ptxas info    :", "").strip() for x in lines[i + 1:i + 3]))
    sass = subprocess.run(["cuobjdump", "-sass", CUBIN], capture_output=True, text=True, check=True).stdout
    ins, on = [], False
    for l in sass.splitlines():
        if "Function :" in l:
            on = KEY in l
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if on and m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    at = {a: i for i, (a, _) in enumerate(ins)}
    bra = re.compile(r"BRA(?:\.U)?(?:\.DIV)?\s+((?:!?U?P\d|UR\d+),\s*)?0x([0-9a-f]+)")
    loops, inner = [], []
    for a, x in ins:
        m = bra.search(x)
        if m and int(m.group(2), 16) < a:
            t = int(m.group(2), 16)
            body = [y for (b, y) in ins if t <= b <= a]
            gathers = sum("LDG.E.U8" in y for y in body), sum("TLD4" in y for y in body)
            if gathers in ((32, 0), (0, 8), (0, 12), (0, 6)) and any("UBLKCP" in y for y in body):  # byte loads, or -DVORS_TEX=1: one gather each
                loops.append((a - t, t, a))
            if gathers == (0, 6) and not any("UBLKCP" in y for y in body):
                inner.append((a - t, t, a))
    _, head, tail = min(loops)
    start, end = at[head], at[tail]
    dist, prev, pq = {start: 0}, {}, [(0, start)]
    while pq:
        d, i = heapq.heappop(pq)
        if d > dist.get(i, 1 << 60) or i == end:
            continue
        x = ins[i][1]
        m = bra.search(x)
        if m:
            succ = [at[int(m.group(2), 16)]] if int(m.group(2), 16) in at else []
            if x.startswith("@") or m.group(1):
                succ.append(i + 1)
        elif x.startswith(("EXIT", "RET")):
            succ = []
        else:
            succ = [i + 1]
        for s in succ:
            if s < len(ins) and dist.get(s, 1 << 60) > d + 1:
                dist[s], prev[s] = d + 1, i
                heapq.heappush(pq, (d + 1, s))
    path, i = [], end
    while i != start:
        path.append(i)
        i = prev[i]
    path.append(start)
    path.reverse()
    ops = collections.Counter()
    for i in path:
        p = ins[i][1].split()
        ops[(p[1] if p[0].startswith("@") else p[0]).split(".")[0]] += 1
    n_words = max(8, sum("TLD4" in ins[i][1] for i in path))  # words (= candidates per lane) per iteration: 8, or 12 tiled
    if inner:  # tiled records: the unrolled half tile (six words) is an inner loop that runs twice per stage
        _, ih, it = min(inner)
        n_in = sum(1 for i in path if ih <= ins[i][0] <= it)
        print(f"inner half-tile loop {ih:#x}..{it:#x}: {(it - ih) // 16 + 1} instructions in the body, {n_in} on the common path = "
              f"{n_in / 6:.3f} per candidate; per stage of 12: {len(path) + n_in} = {(len(path) + n_in) / 12:.3f} per candidate")
        n_words = 6
    print(f"loop {head:#x}..{tail:#x}: {(tail - head) // 16 + 1} instructions in the body, common path {len(path)} "
          f"= {len(path) / n_words:.3f} per candidate ({n_words} per iteration); BRA.DIV on it: {sum('BRA.DIV' in ins[i][1] for i in path)}")
    print(sorted(ops.items(), key=lambda kv: -kv[1]))
    with open("/tmp/common.sass", "w") as f:
        f.write("\n".join(f"{ins[i][0]:05x} {ins[i][1]}" for i in path))


if __name__ == "__main__":
    main()
