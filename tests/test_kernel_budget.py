"""CPU: static budget of the default step kernel, read from the compiled object (no GPU needed).

The const-slot march of csrc/fused2b.cu was tuned on its instruction mix (profiles/r1_sass_static.md):
these checks keep a refactor from silently giving the gains back -- register spills, a register count
that no longer allows 4 CTAs/SM, ring-slot arithmetic creeping back into the march loop."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "py-cubed-sphere_b200", "build", "fused2b.o")
LOG = os.path.join(ROOT, "py-cubed-sphere_b200", "build", "fused2b.cu.ptxas.log")
DEFAULT = "Li160ELi3ELi1ELi%dELi0ELi32EEEv"     # fused2b_kernel<160, 3, 1, MASK, GH = 0, VAR = 32 (F2B_VAR_DEFAULT)>

pytestmark = pytest.mark.skipif(not (os.path.exists(OBJ) and os.path.exists(LOG) and shutil.which("cuobjdump")),
                                reason="needs the in-tree build (python __graft_entry__.py) and cuobjdump")


def test_no_step_kernel_spills_and_default_fits_four_ctas():
    txt = open(LOG).read()
    entries = re.findall(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                         r"(\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt)
    assert len(entries) >= 30
    for name, stack, st, ld, regs in entries:
        # one 8-byte spill slot is tolerated in the SP-L04 / SP-PL07 instantiations; none in the default scheme.
        # The one-kernel flavour (GH = 2, "ELi2ELi32EEEv") spills around the call of its ghost prologue -- once per
        # CTA, outside the march (test_one_kernel_flavour_keeps_the_march_clean looks at the loop itself)
        lim = 128 if "ELi2ELi32EEEv" in name else 16
        assert int(stack) <= max(8, lim) and int(st) <= max(8, lim) and int(ld) <= lim, name
        assert int(regs) * 160 * 4 <= 65536, (name, regs)
    for mask in (0, 1, 2):
        for gh in (0, 1, 2):
            hit = [(int(a), int(b), int(c)) for n, a, b, c, _ in entries if (DEFAULT % mask).replace("ELi0ELi32EEEv", "ELi%dELi32EEEv" % gh) in n]
            # the interior / single-GPU flavour (GH = 0) must not spill at all; the boundary flavour of the
            # multi-GPU split step (GH = 1, in-kernel exchange) and the one-kernel flavour (GH = 2) may keep one 8-byte slot
            assert hit and (hit[0] == (0, 0, 0) if gh == 0 else max(hit[0]) <= (8 if gh == 1 else 128)), (mask, gh, hit)


@pytest.mark.parametrize("mask", [0, 1, 2])
def test_one_kernel_flavour_keeps_the_march_clean(mask):
    """GH = 2: the march loop of the CTAs without ghost cells is the GH = 0 loop -- same instruction count, no
    local-memory traffic (the spills ptxas reports belong to the ghost prologue)."""
    name = (DEFAULT % mask).replace("ELi0ELi32EEEv", "ELi2ELi32EEEv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sass_stats.py"), OBJ, name],
                         capture_output=True, text=True, check=True).stdout
    every = float(re.search(r"every warp: \d+ instructions = ([\d.]+) per row", out).group(1))
    assert every <= (195.0 if mask != 1 else 215.0), out
    ops = out.split("opcodes (every warp):")[1]
    assert "LDL" not in ops and "STL" not in ops, ops


@pytest.mark.parametrize("mask", [0, 2])
def test_default_march_loop_instruction_mix(mask):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sass_stats.py"), OBJ, DEFAULT % mask],
                         capture_output=True, text=True, check=True).stdout
    rows = int(re.search(r"(\d+) rows per trip", out).group(1))
    assert rows == 6                                   # the six-row group of the const-slot march
    common = out.split("elected lane")[0]              # the part every warp executes
    per = {m.group(1): float(m.group(2)) for m in re.finditer(r"^\s+(\S+)\s+\d+\s+([\d.]+) / row", common, re.M)}
    every = float(re.search(r"every warp: \d+ instructions = ([\d.]+) per row", out).group(1))
    assert every <= 195.0, out                         # 232.6 for the shifting-window march, 194.3 at the end of round 1
    issue = float(re.search(r"elected lane of warp 0 \(TMA issue\): \d+ instructions = ([\d.]+) per row", out).group(1))
    assert issue <= 40.0, out                          # 70.2 with bases + byte offset: the other four warps wait for it
    assert per["fp64"] <= 97.0 and per["lds"] <= 28.5 and per["int"] <= 20.0, out
    assert per["bar"] == 2.0
