"""Experiment builds: python tools/buildv.py <name|main> [DEFINE ...] -> exp/lib_<name>.so; prints the register use of k_zstat<double, 20, 0>."""
import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesnmf_b200 import build
name, defs = sys.argv[1], tuple(d for d in sys.argv[2:] )
out = None if name == "main" else f"/root/repo/exp/lib_{name}.so"
import io, contextlib
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    build.build(force=True, verbose=True, defines=defs, out=out)
txt = buf.getvalue().splitlines()
for i, l in enumerate(txt):
    if "k_zstatIdLi20ELi0" in l and "Function properties" in l:
        print(name, txt[i+1].strip(), "|", txt[i+2].strip())
