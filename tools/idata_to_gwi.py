#!/usr/bin/env python
"""Convert the reference's NetCDF-4 / HDF5 catalog (written by
gwinferno.preprocess.data_collection.save_posterior_samples_and_injection_datasets_as_idata, data_collection.py:203-207:
ArviZ InferenceData with groups pe_data.posteriors[event, param, samples] and inj_data.injections[param, injection],
attributes total_generated / analysis_time) into the flattened catalog gwinferno_b200.catalog_io reads.

Runs where ArviZ + xarray exist (NOT in the build container or on the GPU box: untested there):

    python tools/idata_to_gwi.py catalog.h5 catalog.npz        # or catalog.nc (NetCDF classic)
"""
import sys

import numpy as np


def main(src, dst):
    import arviz as az  # noqa: the one place that needs it

    sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
    from gwinferno_b200 import catalog_io

    data = az.from_netcdf(src)
    pe, inj = data.pe_data, data.inj_data
    pedict = {str(k): np.asarray(pe.posteriors.sel(param=k).values, dtype=np.float64) for k in pe.param.values}
    injdict = {str(k): np.asarray(inj.injections.sel(param=k).values, dtype=np.float64) for k in inj.param.values}
    catalog_io.save_catalog(dst, pedict, injdict, float(inj.attrs["total_generated"]), float(inj.attrs["analysis_time"]), events=[str(e) for e in pe["event"].values])
    print(f"{dst}: {len(pedict)} PE parameters x {next(iter(pedict.values())).shape}, {next(iter(injdict.values())).shape[0]} injections")


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    main(sys.argv[1], sys.argv[2])
