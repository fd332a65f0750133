"""The reference's record / file formats for the hot path.  Parsing only.

  demand CSV    `id,from,to,maxWait,maxLoss` per line            pool_n.c:30-54 (readDemand)
  result CSV    `p0,..,p{k-1},d0,..,d{k-1},cost,` per plan        pool_n.c:64-81 (writeResult)
  findpool CSV  same without the cost column                     findpool.c:44-63
  cost.txt      first line n, then n lines of n ints             Simulator.java:512-518 -> solver.py:30-32
  solv_out.txt  n*n lines `0` / `1`, x[n*cab+cust]               solver.py:36-39 -> Simulator.java:306-326
  taxi_demand   `(id,from,to,time,at)` tuples, whitespace separated   gendemand.py:2-20, Simulator.java
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np

REC_W = 9


def _atoi(tok: str) -> int:
    """C atoi: optional blanks, optional sign, leading digits; anything else is 0."""
    t = tok.lstrip(" \t\n\v\f\r")
    i = 1 if t[:1] in "+-" else 0
    j = i
    while j < len(t) and t[j].isdigit():
        j += 1
    return int(t[:j]) if j > i else 0


def read_demand_csv(text: str, max_rows: int | None = None, pad_to: int | None = None) -> np.ndarray:
    """readDemand (pool_n.c:30-54), literally: fgets with a 40-byte buffer (a longer line continues in the next record),
    strtok on ',', atoi on up to five fields; fields that are not given keep the zero of the static array, and so does
    every record the file does not reach.  A blank line is a record (of zeros), exactly as in the reference.
    max_rows = the reference's `linesNumb` (stop after that many records); pad_to: number of rows of the result
    (the reference always searches `rec-number` records, pool_n.c:221,229 -- rows past the end of the file are zeros)."""
    rows = []
    pos, n_text = 0, len(text)
    while pos < n_text and (max_rows is None or len(rows) < max_rows):
        end = text.find("\n", pos, pos + 39)              # fgets: at most 39 characters, stops after a newline
        nxt = end + 1 if end >= 0 else min(pos + 39, n_text)
        line = text[pos:nxt]
        pos = nxt
        toks = [t for t in line.split(",") if t != ""]    # strtok skips empty fields
        r = [0, 0, 0, 0, 0]
        for i, tok in enumerate(toks[:5]):
            r[i] = _atoi(tok)
        rows.append(r)
    if pad_to is not None:
        rows = rows[:pad_to] + [[0, 0, 0, 0, 0]] * max(0, pad_to - len(rows))
    return np.asarray(rows, dtype=np.int32).reshape(-1, 5)


def write_demand_csv(rows) -> str:
    return "".join("%d,%d,%d,%d,%d\n" % tuple(int(v) for v in r) for r in np.asarray(rows).reshape(-1, 5))


def write_result_csv(plans, pool_size: int, with_cost: bool = True) -> str:
    """pool_n.c:72-78 (with_cost) / findpool.c:52-59 (without)."""
    out = []
    for r in np.asarray(plans).reshape(-1, REC_W):
        line = "".join("%d," % int(v) for v in r[: 2 * pool_size])
        out.append(line + ("%d,\n" % int(r[8]) if with_cost else "\n"))
    return "".join(out)


def read_result_csv(text: str, pool_size: int) -> np.ndarray:
    rows = []
    for line in text.splitlines():
        f = [int(v) for v in line.strip().split(",") if v != ""]
        if not f:
            continue
        r = [0] * REC_W
        r[: 2 * pool_size] = f[: 2 * pool_size]
        if len(f) > 2 * pool_size:
            r[8] = f[2 * pool_size]
        rows.append(r)
    return np.asarray(rows, dtype=np.int32).reshape(-1, REC_W)


def write_cost_txt(cost) -> str:
    """Simulator.java:512-518: `n\\n` then n rows, every value followed by one blank."""
    c = np.asarray(cost)
    n = c.shape[0]
    return "%d\n" % n + "".join("".join("%d " % int(v) for v in row) + "\n" for row in c)


def read_cost_txt(text: str) -> Tuple[int, np.ndarray]:
    """solver.py:30-32."""
    lines = text.splitlines()
    n = int(lines[0])
    cost = [[int(x) for x in line.split()] for line in lines[1:] if line.strip()]
    return n, np.asarray(cost, dtype=np.int32).reshape(n, n)


def write_solv_out(x: Sequence[int]) -> str:
    """solver.py:36-39."""
    return "".join("%d\n" % int(v) for v in x)


def read_solv_out(text: str, n: int) -> np.ndarray:
    """Simulator.java:306-326."""
    vals = [int(v) for v in text.split()]
    if len(vals) != n * n:
        raise ValueError("solver output has %d values, expected %d" % (len(vals), n * n))
    return np.asarray(vals, dtype=np.int32)


def read_taxi_demand(text: str) -> List[Tuple[int, int, int, int, int]]:
    """simulations/taxi_demand.txt: whitespace separated `(id,from,to,time,at)`."""
    out = []
    for tok in text.split():
        a = tok.strip("()").split(",")
        out.append(tuple(int(v) for v in a))
    return out
