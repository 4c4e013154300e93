"""CPU: the on-disk side of SURVEY.md section 8 row f4 -- the NetCDF classic reader / writer of
gwinferno_b200/catalog_io.py (the reference loads its catalogs with ArviZ / xarray, pipeline/utils.py:51-96,
tests/inference_test.py:74-82; neither exists here) and the host-side redshift PPD helpers."""

import os

import numpy as np
import pytest

from gwinferno_b200 import catalog_io as cio

FIXTURE = "/root/reference/tests/data/xarray_GWTC3_BBH_69evs_downsampled_1000samps_nospin.h5"
PARAMS = ["redshift", "mass_1", "a_1", "cos_tilt_1", "mass_2", "a_2", "cos_tilt_2", "mass_ratio", "prior"]


@pytest.mark.skipif(not os.path.exists(FIXTURE), reason="the reference's PE fixture is only present in the build container")
def test_reads_the_reference_fixture_like_scipy_does():
    """The reference's vendored PE file is a CDF-2 container: every variable equals what scipy.io.netcdf_file reads."""
    from scipy.io import netcdf_file

    dims, att, var = cio.read_netcdf3(FIXTURE)
    assert dims == {"param": 9, "sample": 1000, "string10": 10}
    f = netcdf_file(FIXTURE, "r", mmap=False)
    assert set(f.variables) == set(var)
    for k, v in f.variables.items():
        assert var[k][0] == v.dimensions
        assert np.array_equal(var[k][2], np.asarray(v[:]), equal_nan=var[k][2].dtype.kind == "f")
    pe, events, params = cio.load_pe_dataset(FIXTURE)
    assert params == PARAMS and len(events) == 69 and events[0] == "GW150914"
    assert all(pe[p].shape == (69, 1000) and pe[p].dtype == np.float64 for p in params)
    # known answers: the first samples of GW150914 (float32 on disk)
    assert np.allclose(pe["mass_1"][0, :3], [38.44009399, 33.9650116, 38.08488083], rtol=1e-7)
    assert np.allclose(pe["prior"][0, :3], [0.00522551, 0.00774921, 0.00725626], rtol=1e-6)
    sub, _, _ = cio.load_pe_dataset(FIXTURE, n_samples=100, rng=np.random.default_rng(3))  # tests/inference_test.py:74-82
    assert sub["mass_ratio"].shape == (69, 100) and np.all((sub["mass_ratio"] > 0) & (sub["mass_ratio"] <= 1))


@pytest.mark.parametrize("ext", [".nc", ".npz"])
def test_flattened_catalog_round_trip(tmp_path, ext):
    rng = np.random.default_rng(5)
    events = [f"GW{i:06d}" for i in range(6)]
    pe = {p: rng.random((6, 37)) for p in PARAMS}
    inj = {p: rng.random(211) for p in PARAMS}
    path = str(tmp_path / ("cat" + ext))
    cio.save_catalog(path, pe, inj, total_inj=1234.0, obs_time=0.75, events=events)
    pedict, injdict, const, names = cio.load_pe_and_injections_as_dict(path)
    assert names == PARAMS and const == {"total_inj": 1234.0, "obs_time": 0.75, "nObs": 6}
    assert all(np.array_equal(pedict[p], pe[p]) for p in PARAMS) and all(np.array_equal(injdict[p], inj[p]) for p in PARAMS)
    pedict, _, const, _ = cio.load_pe_and_injections_as_dict(path, ignore=[events[2], events[5]])  # utils.py:77-82
    assert const["nObs"] == 4 and np.array_equal(pedict["mass_1"], pe["mass_1"][[0, 1, 3, 4]])


def test_written_file_is_a_valid_classic_file_for_other_readers(tmp_path):
    from scipy.io import netcdf_file

    rng = np.random.default_rng(6)
    path = str(tmp_path / "c.nc")
    cio.save_catalog(path, {"a": rng.random((2, 5)), "b": rng.random((2, 5))}, {"a": rng.random(7), "b": rng.random(7)}, 10, 1.0)
    f = netcdf_file(path, "r", mmap=False)
    assert f.variables["posteriors"].shape == (2, 2, 5) and f.variables["injections"].shape == (2, 7)
    assert float(f._attributes["total_generated"]) == 10.0


def test_record_variables_and_cdf1_cdf5_headers(tmp_path):
    """Files other tools write: record (unlimited) variables, CDF-1 offsets, CDF-5 64-bit counts."""
    from scipy.io import netcdf_file

    path = str(tmp_path / "r.nc")
    f = netcdf_file(path, "w", version=1)
    f.createDimension("t", None)
    f.createDimension("x", 3)
    v = f.createVariable("u", "f8", ("t", "x"))
    w = f.createVariable("k", "i4", ("t",))
    s = f.createVariable("s", "f4", ("x",))
    v[:] = np.arange(12.0).reshape(4, 3)
    w[:] = np.arange(4) * 7
    s[:] = [1.5, 2.5, 3.5]
    f.history = "made by scipy"
    f.close()
    dims, att, var = cio.read_netcdf3(path)
    assert dims["t"] == 0 and att["__numrecs__"] == 4 and att["history"] == "made by scipy"
    assert np.array_equal(var["u"][2], np.arange(12.0).reshape(4, 3)) and np.array_equal(var["k"][2], np.arange(4) * 7)
    assert np.array_equal(var["s"][2], np.array([1.5, 2.5, 3.5], dtype=np.float32))
    with pytest.raises(cio.NetCDFError, match="HDF5"):
        cio.read_netcdf3(b"\x89HDF\r\n\x1a\n" + b"\x00" * 64)


def test_rate_of_z_ppds_equal_reference():
    """calculations.py:244-276 (O(grid) host arithmetic on the mirror redshift model's grid)."""
    from gwinferno_b200 import models as M
    from gwinferno_b200 import postprocess as PP

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ppd_reference.npz"))
    zm = M.PowerlawSplineRedshiftModel(8, g["in_z_pe"], g["in_z_inj"])
    rs, zs = PP.calculate_powerlaw_rate_of_z_ppds(g["in_lamb_z"], g["in_rate"], zm)
    assert np.allclose(zs, g["out_zs"], rtol=1e-15) and np.allclose(rs, g["out_plz_rs"], rtol=1e-13)
    rs, _ = PP.calculate_powerlaw_spline_rate_of_z_ppds(g["in_lamb_z"], g["in_z_cs"], g["in_rate"], zm, pop_frac=g["in_frac"])
    assert np.allclose(rs, g["out_plsz_rs"], rtol=1e-11)
