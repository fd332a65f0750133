"""CPU suite: the reference's file formats round-trip and match the compiled reference's files."""
import numpy as np

from oracle import gen_inputs as g
from taxidispatcher_b200 import formats
from conftest import load_golden


def test_demand_csv_matches_reference_input_format():
    dem = g.pool_demand(50, seed=3)
    txt = formats.write_demand_csv(dem)
    assert txt == g.demand_csv(dem)
    assert np.array_equal(formats.read_demand_csv(txt), dem)
    assert np.array_equal(formats.read_demand_csv(txt, 7), dem[:7])
    assert all(len(line) < 39 for line in txt.splitlines())   # pool_n.c:33 char line[40]


def test_demand_csv_read_demand_quirks():
    """readDemand (pool_n.c:30-54) read literally: a blank line is a record of zeros, missing fields stay zero, records
    the file does not reach stay zero (the search still runs over `rec-number` customers), atoi swallows junk, and a
    line longer than the 40-byte fgets buffer spills into the next record."""
    txt = "0,5,7,3,1\n\n2,9\nx,4,-6,+2,1abc,99\n"
    got = formats.read_demand_csv(txt, 6, pad_to=6)
    assert got.tolist() == [[0, 5, 7, 3, 1], [0, 0, 0, 0, 0], [2, 9, 0, 0, 0], [0, 4, -6, 2, 1], [0, 0, 0, 0, 0], [0, 0, 0, 0, 0]]
    assert formats.read_demand_csv(txt, 2).tolist() == [[0, 5, 7, 3, 1], [0, 0, 0, 0, 0]]          # linesNumb stops the loop
    long_line = "1," + "0" * 40 + "7,8,9,10\n3,4,5,6,7\n"
    got = formats.read_demand_csv(long_line)
    assert got[0, 0] == 1 and len(got) == 3 and got[2].tolist() == [3, 4, 5, 6, 7]                  # first 39 bytes, the rest, next line
    assert formats.read_demand_csv("", pad_to=3).tolist() == [[0] * 5] * 3


def test_result_csv_round_trip():
    case = load_golden("pool_small.json")[5]
    k = case["pool_size"]
    plans = np.array(case["shards"][0]["plans"], dtype=np.int32).reshape(-1, 9)
    txt = formats.write_result_csv(plans, k)
    assert txt.splitlines()[0].endswith(",") and txt.count("\n") == len(plans)
    assert np.array_equal(formats.read_result_csv(txt, k), plans)
    no_cost = formats.write_result_csv(plans, k, with_cost=False)          # findpool.c:52-59
    back = formats.read_result_csv(no_cost, k)
    assert np.array_equal(back[:, : 2 * k], plans[:, : 2 * k]) and (back[:, 8] == 0).all()


def test_cost_and_solver_files():
    c = g.config1a(7)
    txt = formats.write_cost_txt(c)
    assert txt.splitlines()[0] == "7" and txt.splitlines()[1].endswith(" ")
    n, back = formats.read_cost_txt(txt)
    assert n == 7 and np.array_equal(back, c)
    x = np.eye(7, dtype=np.uint8).reshape(-1)
    out = formats.write_solv_out(x)
    assert out.count("\n") == 49
    assert np.array_equal(formats.read_solv_out(out, 7), x)


def test_taxi_demand_tokens():
    rows = formats.read_taxi_demand("(0,37,40,0,7)\n(1,10,6,0,0) (2,30,31,0,0)")
    assert rows == [(0, 37, 40, 0, 7), (1, 10, 6, 0, 0), (2, 30, 31, 0, 0)]
