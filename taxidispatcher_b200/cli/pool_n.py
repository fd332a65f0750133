"""`pool_n`-compatible CLI (pool_n.c:209-238):

    python -m taxidispatcher_b200.cli.pool_n <pool-size> <thread 0..7> <demand-file> <rec-number> <output-file>

Reads the demand CSV, runs ONE logical shard on the GPU, writes the result CSV in the reference's
format and drops `out<thread>.flg` in the current directory (pool_n.c:56-62), so that an unmodified
findpool.c can drive it.  Errors print a message and exit(1) like the reference (pool_n.c:36-39).
"""
from __future__ import annotations

import sys

import numpy as np


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 5:
        print("Usage: cmd pool-size thread-number demand-file-name rec-number output-file", end="")
        return 1
    pool_size, thread, fname, rec_number, out_name = int(argv[0]), int(argv[1]), argv[2], int(argv[3]), argv[4]
    from .. import dispatch, formats
    try:
        text = open(fname).read()
    except OSError:
        print("Opening file %s failed" % fname)
        return 1
    demand = formats.read_demand_csv(text, rec_number, pad_to=rec_number)   # ordersCount = rec-number whatever the file holds (pool_n.c:219-221)
    n_stands = 51                                                      # pool_n.c:15 MAX_STAND
    idx = np.arange(n_stands, dtype=np.int32)
    dist = np.abs(idx[:, None] - idx[None, :]).astype(np.int32)       # pool_n.c:179-185 setCosts
    plans, _ = dispatch.find_pool(demand, dist, pool_size, thread, 8)
    try:
        with open(out_name, "w") as f:
            f.write(formats.write_result_csv(plans, pool_size))
    except OSError:
        print("Opening file %s failed" % out_name)
        return 1
    with open("out%d.flg" % thread, "w") as f:                         # pool_n.c:56-62 touchFile
        f.write("%d" % thread)
    return 0


if __name__ == "__main__":
    sys.exit(main())
