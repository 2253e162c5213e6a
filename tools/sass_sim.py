"""Static timing of the basic blocks of a kernel for ONE warp alone on a scheduler, from the
control words ptxas wrote into the SASS (stall count, scoreboard set / wait) and measured latencies
of the variable-latency operations (tools/ubench/lat.cu): what the schedule costs if nothing else
runs, block by block, and which scoreboard waits it is made of.  No GPU needed.

usage: python tools/sass_sim.py <sass from cuobjdump -sass> [min_instructions]"""
import collections
import re
import sys

LAT = {"MUFU": 18, "SHFL": 25, "LDS": 29, "LDG": 40, "LDC": 10, "LDCU": 10, "S2R": 20, "LDL": 30, "STS": 10,
       "STG": 10, "STL": 10, "VOTE": 12, "VOTEU": 12, "R2UR": 8, "S2UR": 20, "F2F": 12, "I2F": 12, "F2I": 12,
       "I2FP": 12, "F2FP": 12, "DFMA": 12, "DMUL": 12, "DADD": 12, "REDUX": 20, "ATOMG": 60, "ATOMS": 40,
       "BAR": 10, "POPC": 10, "FLO": 10, "BREV": 10, "IMAD.WIDE": 6, "CCTL": 10, "MATCH": 20, "DSETP": 12}
MUFU_PIPE = 8  # cycles of the MUFU unit per warp instruction


def parse(path):
    ins = []
    lines = open(path).read().split("\n")
    i = 0
    while i < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            hi = int(m2.group(1), 16) if m2 else 0
            addr = int(m.group(1), 16)
            text = m.group(2).strip()
            t = text.split()
            op = t[1] if t[0].startswith("@") else t[0]
            c = hi >> 41
            ins.append(dict(addr=addr, text=text, op=op, stall=c & 0xF, wbar=(c >> 5) & 7, rbar=(c >> 8) & 7,
                            wait=(c >> 11) & 0x3F))
            i += 2
        else:
            i += 1
    return ins


def blocks(ins):
    targets = set()
    for x in ins:
        if x["op"].split(".")[0] in ("BRA", "BSSY", "CALL", "BRX", "JMP"):
            for m in re.finditer(r"0x([0-9a-f]+)", x["text"]):
                targets.add(int(m.group(1), 16))
    out, cur = [], []
    for x in ins:
        if x["addr"] in targets and cur:
            out.append(cur)
            cur = []
        cur.append(x)
        if x["op"].split(".")[0] in ("BRA", "EXIT", "RET", "BRX", "JMP", "CALL", "BREAK"):
            # a predicated branch ends the block too: ptxas does not schedule across it
            out.append(cur)
            cur = []
    if cur:
        out.append(cur)
    return out


def simulate(blk):
    t = 0
    sb = [0] * 6          # time at which each scoreboard clears
    mufu_free = 0
    waits = collections.Counter()
    for x in blk:
        ready = t
        for b in range(6):
            if x["wait"] >> b & 1 and sb[b] > ready:
                ready = sb[b]
        if ready > t:
            waits[x["op"].split(".")[0]] += ready - t
            t = ready
        base = x["op"].split(".")[0]
        lat = LAT.get(x["op"]) or LAT.get(base) or 6
        issue = t
        if base == "MUFU":
            issue = max(t, mufu_free)
            mufu_free = issue + MUFU_PIPE
        done = issue + lat
        if x["wbar"] != 7:
            sb[x["wbar"]] = max(sb[x["wbar"]], done)
        if x["rbar"] != 7:
            sb[x["rbar"]] = max(sb[x["rbar"]], issue + 4)
        t += max(1, x["stall"])
    return t, waits


if __name__ == "__main__":
    ins = parse(sys.argv[1])
    minn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    print(f"{len(ins)} instructions")
    for blk in blocks(ins):
        if len(blk) < minn:
            continue
        cyc, waits = simulate(blk)
        stalls = sum(max(1, x["stall"]) for x in blk)
        ops = collections.Counter(x["op"].split(".")[0] for x in blk)
        print(f"block {blk[0]['addr']:#07x}..{blk[-1]['addr']:#07x}: {len(blk):5d} instr, stall counts {stalls:5d}, "
              f"simulated {cyc:5d} cycles; scoreboard waits {sum(waits.values())} "
              f"({', '.join(f'{k} {v}' for k, v in waits.most_common(5))}); MUFU {ops['MUFU']}, SHFL {ops['SHFL']}, LDS {ops['LDS']}")
