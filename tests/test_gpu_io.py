"""GPU: host <-> device I/O formats against the reference's goldens
(spec/n_array_spec.cr:520-558)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D, io


def test_json_yaml_goldens():
    stock = D.from_host(np.array([[0, 1, 2], [3, 4, 5]], np.int32))
    assert io.to_json(stock) == '{"shape":[2,3],"elements":[0,1,2,3,4,5]}'
    assert io.to_json(D.fill([0], 0, np.int32)) == '{"shape":[0],"elements":[]}'
    assert io.from_json('{"shape":[2,3],"elements":[0,1,2,3,4,5]}', np.int32).equals(stock)
    e = io.from_json('{"shape":[0],"elements":[]}', np.int32)
    assert e.shape == [0] and e.size == 0
    assert io.to_yaml(stock) == "---\nshape: [2, 3]\nelements: [0, 1, 2, 3, 4, 5]\n"
    assert io.to_yaml(D.fill([0], 0, np.int32)) == "---\nshape: [0]\nelements: []\n"
    assert io.from_yaml("---\nshape: [2, 3]\nelements: [0, 1, 2, 3, 4, 5]\n", np.int32).equals(stock)
    assert io.from_yaml("---\nshape: [0]\nelements: []\n", np.int32).shape == [0]
    with pytest.raises(ph.ShapeError):
        io.from_json('{"shape":[2,2],"elements":[1,2,3]}', np.int32)
    with pytest.raises(ValueError):
        io.from_json('{"shape":[2,2]}', np.int32)


def test_float_round_trip_and_binary_dump(tmp_path):
    rs = np.random.RandomState(0)
    a = rs.rand(7, 5, 3)
    d = D.from_host(a)
    assert io.from_json(io.to_json(d), np.float64).equals(d)            # repr round-trips f64 exactly
    assert io.from_yaml(io.to_yaml(d), np.float64).equals(d)
    big = D.from_host(rs.rand(300, 200).astype(np.float32))
    path = str(tmp_path / "state.phbin")
    io.dump(big, path)
    back = io.load(path)
    assert back.dtype == np.float32 and back.equals(big)
    # checkpoint / resume of a heat run: 3 + 4 steps == 7 steps, bit for bit
    from ph_core_b200 import heat
    s = D.from_host((rs.rand(12, 20, 64) * 100).astype(np.float32))
    whole = heat.simulate(s.clone(), 0.1, 7)
    part = heat.simulate(s.clone(), 0.1, 3)
    io.dump(part, path)
    resumed = heat.simulate(io.load(path), 0.1, 4)
    assert resumed.equals(whole)
