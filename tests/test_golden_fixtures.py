"""The committed fixtures are what tests/golden/make_golden.py extracts from the reference checkout -- checked wherever that checkout
exists (the build container); skipped on the GPU box, where only the fixtures travel."""
import json
import os
import sys

import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "test", "data")), reason="reference checkout not present")
def test_committed_fixtures_match_the_reference_files():
    sys.path.insert(0, GOLDEN)
    import make_golden

    for name, make in (("reference_histories.json", make_golden.histories), ("reference_vectors.json", make_golden.vectors),
                       ("sparta_couette.json", make_golden.sparta)):
        fresh = json.loads(json.dumps(make()))
        committed = json.load(open(os.path.join(GOLDEN, name)))
        assert fresh == committed, name


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "test", "data")), reason="reference checkout not present")
def test_hdf5_reader_on_a_golden_file():
    """hdf5_min.py on test/data/2species_seed1234.nc: dataset names, shapes, dtypes and the first values the reference wrote."""
    sys.path.insert(0, GOLDEN)
    from hdf5_min import H5File

    f = H5File(os.path.join(REF, "test", "data", "2species_seed1234.nc"))
    assert {"timestep", "np", "ndens", "v", "T", "moments", "moment_powers"} <= set(f.datasets)
    T, npart, ts = f.read("T"), f.read("np"), f.read("timestep")
    assert T.shape == (801, 2, 1) and str(T.dtype) == "float64" and f.read("v").shape == (801, 2, 1, 3)
    assert list(ts[:3]) == [0.0, 1.0, 2.0] and ts[-1] == 800.0
    assert list(npart[0, :, 0]) == [400.0, 4000.0] and list(f.read("ndens")[0, :, 0]) == [2e15, 2e16]
    assert abs(T[0, 0, 0] - 3014.32157736) < 1e-8 and abs(T[800, 1, 0] - 598.25911168) < 1e-8
