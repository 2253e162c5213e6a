"""What the built library contains, read from its SASS with the CUDA binary tools (no GPU): which
arithmetic each mode really runs, the register budgets the launch shapes rely on, and the static
schedule of the two-lanes-per-cell sub-step (tools/sass_sim.py) as a regression guard."""
import collections
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hybrid9_b200", "libh9gpu.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)),
                                reason="needs the built library and cuobjdump")


@pytest.fixture(scope="module")
def sass():
    out = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fns, cur = collections.OrderedDict(), None
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fns[cur] = []
        elif cur is not None:
            fns[cur].append(line)
    return fns


@pytest.fixture(scope="module")
def usage():
    out = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.split("\n"):
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and cur:
            res[cur] = tuple(int(v) for v in m.groups())
            cur = None
    return res


def opcodes(lines):
    c = collections.Counter()
    for line in lines:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            c[m.group(1)] += 1
            c[m.group(1).split(".")[0]] += 0
    return c


def pick(fns, *needles):
    hits = [k for k in fns if all(n in k for n in needles)]
    assert hits, needles
    return hits


def test_only_sm_100a_code(sass):
    out = subprocess.run([CUOBJDUMP, "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_fast_kernels_run_on_the_mufu_unit_without_ieee_division(sass):
    """H9_MATH_FAST: pow = ex2.approx(b * lg2.approx(a)), 1/x = rcp.approx; no double precision;
    the only IEEE divisions (FCHK-guarded sequences) are the 36-40 that fill the per-cell constant
    table once per launch (CellTable::init / PairTable::init), against 327 in an exact kernel."""
    for k in pick(sass, "days_kernel_fast") + pick(sass, "days_kernel_pair"):
        c = opcodes(sass[k])
        assert c["MUFU.EX2"] >= 40 and c["MUFU.LG2"] >= 40 and c["MUFU.RCP"] >= 10, (k, c["MUFU.EX2"], c["MUFU.LG2"])
        assert c["FCHK"] <= 40, (k, c["FCHK"])
        assert sum(v for o, v in c.items() if o.startswith(("DFMA", "DMUL", "DADD"))) == 0, k


def test_exact_kernels_use_ieee_division_and_the_portable_double_kernels(sass):
    """H9_MATH_EXACT: IEEE division (the FCHK-guarded sequence), pow/exp/log through the FP64
    kernels (DFMA), no approximate exponentials or logarithms, no fused multiply-add contraction
    of float arithmetic beyond the division and the explicit fmaf of day_setup."""
    for k in pick(sass, "days_kernel", "MathExact"):
        c = opcodes(sass[k])
        assert c["FCHK"] > 100 and c["DFMA"] >= 10, (k, c["FCHK"], c["DFMA"])
        assert c["MUFU.EX2"] == 0 and c["MUFU.LG2"] == 0, k
        calls = sum(v for o, v in c.items() if o.startswith("CALL"))
        assert calls > 300, (k, calls)  # pow/exp/log and the division slow path are calls, not inlined


def test_register_budgets_of_the_launch_shapes(usage):
    """<BLOCK,1> may use every register; the <64,8>/<128,4>/<32,16> builds must fit 16 warps per
    SM (128 registers) -- the one-wave launch of the 0.5 deg grid depends on it -- and no fast
    kernel may touch local memory beyond a small frame."""
    seen = 0
    for k, (reg, stack, shared, local) in usage.items():
        if "days_kernel" not in k:
            continue
        seen += 1
        capped = re.search(r"Li(32)ELi16E|Li(64)ELi8E|Li(128)ELi4E|Li512ELi1E", k) is not None
        assert reg <= (128 if capped else 255), (k, reg)
        assert local == 0, k
        if "MathExact" not in k:
            assert stack <= 160, (k, stack)
    assert seen >= 13, seen


def test_static_schedule_of_the_two_lane_substep(sass, tmp_path):
    """tools/sass_sim.py on the built kernel: the all-deep and the general sub-step are one basic
    block each (branch-free by design) and their static cost for a lone warp stays where the
    profiles record it (955 / 1,249 cycles; 1,048 / 1,367 before the end of the tail was
    reworked)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_sim

    k = pick(sass, "days_kernel_pairILi128")[0]
    p = tmp_path / "pair.sass"
    p.write_text("\n".join(sass[k]))
    ins = sass_sim.parse(str(p))
    big = sorted((b for b in sass_sim.blocks(ins) if len(b) >= 400), key=len)
    assert len(big) == 2, [len(b) for b in sass_sim.blocks(ins) if len(b) > 100]
    deep, general = (sass_sim.simulate(b)[0] for b in big)
    assert 500 <= len(big[0]) <= 700 and 700 <= len(big[1]) <= 900, [len(b) for b in big]
    assert deep <= 1010 and general <= 1320, (deep, general)
