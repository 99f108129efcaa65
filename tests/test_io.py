"""Loaders of the experiments' data formats (asvgp_b200/io.py): round trips through temporary files on the CPU."""
import numpy as np
import pytest


def test_text_loader_reads_the_snelson_layout(tmp_path):
    from asvgp_b200 import io

    x = np.random.default_rng(0).uniform(0, 6, 200)
    p = tmp_path / "train_inputs"
    np.savetxt(p, x)
    t = io.load_text(p)
    assert t.dtype.is_floating_point and t.shape == (200,)
    np.testing.assert_allclose(t.numpy(), x, rtol=1e-15)


def test_pickle_loader_drops_missing_rows_and_rescales(tmp_path):
    import pandas as pd

    from asvgp_b200 import io

    rng = np.random.default_rng(1)
    df = pd.DataFrame({"Date_seq": np.arange(50.0) * 60, "Global_active_power": rng.uniform(0, 5, 50), "other": 1})
    df.loc[[3, 17], "Global_active_power"] = np.nan
    p = tmp_path / "frame.pkl"
    df.to_pickle(p)
    X, y = io.load_pickle(p, "Date_seq", "Global_active_power", rescale_x_to=1000)
    assert X.shape == (48, 1) and y.shape == (48, 1)
    assert X.min().item() == 0.0 and abs(X.max().item() - 1000.0) < 1e-9
    assert np.isfinite(y.numpy()).all()


def test_netcdf_loader_masks_fill_values_and_bbox(tmp_path):
    from scipy.io import netcdf_file

    from asvgp_b200 import io

    p = str(tmp_path / "ssh.nc")
    ny, nx = 12, 15
    lon2, lat2 = np.meshgrid(np.linspace(-80, -25, nx), np.linspace(15, 55, ny))
    ssh = np.sin(lon2 / 9) * np.cos(lat2 / 7)
    ssh[2, 3] = 1e20
    with netcdf_file(p, "w") as nc:
        nc.createDimension("t", 1); nc.createDimension("y", ny); nc.createDimension("x", nx)
        v = nc.createVariable("sossheig", "f8", ("t", "y", "x")); v[0] = ssh; v._FillValue = 1e20
        nc.createVariable("nav_lon", "f8", ("y", "x"))[:] = lon2
        nc.createVariable("nav_lat", "f8", ("y", "x"))[:] = lat2
    X, y = io.load_netcdf(p, "sossheig", bbox=(-75, -30, 20, 50))
    inside = (lon2 > -75) & (lon2 < -30) & (lat2 > 20) & (lat2 < 50)
    inside[2, 3] = False
    assert X.shape == (inside.sum(), 2) and y.shape == (inside.sum(), 1)
    np.testing.assert_allclose(y.numpy().ravel(), ssh[inside], rtol=1e-15)
    (Xtr, ytr), (Xte, yte) = io.train_test_split(X, y, 40, 10, seed=5)
    assert Xtr.shape == (40, 2) and yte.shape == (10, 1)
    rows = {tuple(r) for r in Xtr.numpy()} & {tuple(r) for r in Xte.numpy()}
    assert not rows
    with pytest.raises(ValueError):
        io.load_netcdf(str(tmp_path / "nope.nc"), "sossheig")
