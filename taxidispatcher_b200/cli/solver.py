"""`solver.py`-compatible shim (solver.py:30-39): cost.txt -> solv_out.txt.

    python -m taxidispatcher_b200.cli.solver [cost.txt] [solv_out.txt]

The reference hard-codes c:\\Users\\dell\\TAXI\\out\\cost.txt / solv_out.txt; paths are arguments here."""
from __future__ import annotations

import sys


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    cost_path = argv[0] if len(argv) > 0 else "cost.txt"
    out_path = argv[1] if len(argv) > 1 else "solv_out.txt"
    from .. import dispatch, formats
    nn, cost = formats.read_cost_txt(open(cost_path).read())
    x = dispatch.solve(nn, cost)
    if nn == 0:
        x = []
    with open(out_path, "w") as f:
        f.write(formats.write_solv_out(x))
    return 0


if __name__ == "__main__":
    sys.exit(main())
