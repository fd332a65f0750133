"""`findpool`-compatible CLI (findpool.c:122-176) without the process fan-out:

    python -m taxidispatcher_b200.cli.findpool <pool-size> <demand-file> <rec-number> <output-file>

All 8 logical shards run in one device call, are merged in shard order and written without the
cost column (findpool.c:44-63)."""
from __future__ import annotations

import sys

import numpy as np


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 4:
        print("Usage: findpool pool-size demand-file-name rec-number output-file", end="")
        return 1
    pool_size, fname, rec_number, out_name = int(argv[0]), argv[1], int(argv[2]), argv[3]
    from .. import dispatch, formats
    try:
        text = open(fname).read()
    except OSError:
        print("Opening file %s failed" % fname)
        return 1
    demand = formats.read_demand_csv(text, rec_number, pad_to=rec_number)   # ordersCount = rec-number whatever the file holds (pool_n.c:219-221)
    idx = np.arange(51, dtype=np.int32)
    dist = np.abs(idx[:, None] - idx[None, :]).astype(np.int32)
    plans, _ = dispatch.find_pool_all(demand, dist, pool_size, 8)
    with open(out_name, "w") as f:
        f.write(formats.write_result_csv(plans, pool_size, with_cost=False))
    return 0


if __name__ == "__main__":
    sys.exit(main())
