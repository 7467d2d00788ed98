"""CPU: the labelled-array stand-in (seaduck_b200/minixr.py) behaves like xarray on the part of its surface the
reference touches -- the goldens are made by the unmodified reference running on top of it, so its
semantics are part of what pins the oracle."""

import numpy as np
import pytest

from seaduck_b200 import minixr as xr


@pytest.fixture
def ds():
    rng = np.random.default_rng(0)
    return xr.Dataset(
        coords=dict(Z=(["Z"], -np.arange(1.0, 4.0)), Zl=(["Zl"], -np.arange(0.0, 3.0)), Y=(["Y"], np.arange(4)), X=(["X"], np.arange(5))),
        data_vars=dict(
            T=(["Z", "Y", "X"], rng.normal(size=(3, 4, 5))),
            W=(["Zl", "Y", "X"], rng.normal(size=(3, 4, 5))),
            rA=(["Y", "X"], rng.uniform(1, 2, size=(4, 5))),
            drF=(["Z"], np.array([10.0, 20.0, 30.0])),
            odd=(["X", "Z", "Y"], rng.normal(size=(5, 3, 4))),
        ),
    )


def test_variables_dims_and_attribute_access(ds):
    assert set(ds.data_vars) == {"T", "W", "rA", "drF", "odd"}
    assert {"Z", "Zl", "Y", "X"} <= set(ds.variables)
    assert ds["T"].dims == ("Z", "Y", "X") and ds.T.shape == (3, 4, 5) and ds["T"].ndim == 3
    assert "T" in ds and "nope" not in ds
    with pytest.raises(KeyError):
        ds["nope"]
    sub = ds[["T", "rA"]]
    assert "T" in sub.data_vars and "W" not in sub.data_vars


def test_transpose_reorders_named_dims(ds):
    t = ds.transpose("time", "Z", "Zl", "face", "Y", "X", missing_dims="ignore")  # ocedata.py:165
    assert t["odd"].dims == ("Z", "Y", "X")
    np.testing.assert_array_equal(np.asarray(t["odd"]), np.transpose(np.asarray(ds["odd"]), (1, 2, 0)))
    assert t["T"].dims == ("Z", "Y", "X")
    with pytest.raises(ValueError):
        ds["T"].transpose("time", "Z", "Y", "X")
    e = ds["odd"].transpose("Y", ...)
    assert e.dims == ("Y", "X", "Z")


def test_isel_and_numpy_style_indexing(ds):
    top = ds.isel(Zl=slice(1))  # ocedata.py:202
    assert top["W"].shape == (1, 4, 5) and top["T"].shape == (3, 4, 5)
    np.testing.assert_array_equal(np.asarray(top["W"]), np.asarray(ds["W"])[:1])
    one = ds["T"][1]
    assert one.dims == ("Y", "X")  # an integer index drops the dimension
    np.testing.assert_array_equal(np.asarray(ds["T"][:, 2:4]), np.asarray(ds["T"])[:, 2:4])
    assert float(ds["drF"][1]) == 20.0 and len(ds["drF"]) == 3


def test_label_based_assignment(ds):
    w = ds["W"]
    w.loc[{"Zl": 0}] = 0  # lagrangian.py:155: by coordinate value, not by position
    assert (np.asarray(ds["W"])[0] == 0).all() and (np.asarray(ds["W"])[1] != 0).any()
    w.loc[{"Zl": -2.0}] = 7
    assert (np.asarray(ds["W"])[2] == 7).all()
    with pytest.raises(KeyError):
        w.loc[{"Zl": 5.0}] = 1


def test_named_dimension_broadcasting(ds):
    vol = ds["drF"] * ds["rA"]  # ocedata.py:303-316: dims of the left operand first, then the new ones
    assert vol.dims == ("Z", "Y", "X")
    np.testing.assert_array_equal(np.asarray(vol), np.asarray(ds["drF"])[:, None, None] * np.asarray(ds["rA"])[None])
    vol2 = ds["rA"] * ds["drF"]
    assert vol2.dims == ("Y", "X", "Z")
    np.testing.assert_array_equal(np.asarray(vol2), np.transpose(np.asarray(vol), (1, 2, 0)))
    np.testing.assert_array_equal(np.asarray(ds["T"] - 1), np.asarray(ds["T"]) - 1)
    np.testing.assert_array_equal(np.asarray(2 * ds["T"]), 2 * np.asarray(ds["T"]))


def test_fillna_min_max_astype(ds):
    a = ds["T"].copy()
    a.values[0, 0, 0] = np.nan
    assert np.isnan(np.asarray(a)).sum() == 1 and not np.isnan(np.asarray(a.fillna(0))).any()
    assert float(ds["Z"].min()) == -3.0 and float(ds["Z"].max()) == -1.0
    assert ds["rA"].astype("float32").dtype == np.float32


def test_zeros_like_concat_merge(ds):
    top = ds.isel(Zl=slice(1))
    z = xr.zeros_like(top)
    assert (np.asarray(z["W"]) == 0).all() and z["W"].shape == (1, 4, 5)
    both = xr.concat([ds[["W"]], z[["W"]]], dim="Zl")  # the zero bottom level appended to W (ocedata.py:202-209)
    assert both["W"].shape == (4, 4, 5) and (np.asarray(both["W"])[-1] == 0).all()
    m = xr.merge([ds[["T"]], both])
    assert m["T"].shape == (3, 4, 5) and m["W"].shape == (4, 4, 5)


def test_setitem_accepts_arrays_and_dataarrays(ds):
    with pytest.raises(ValueError):  # like xarray: a bare 2-D array carries no dimension names
        ds["rA"] = np.ones((4, 5))
    ds["rA"] = (["Y", "X"], np.ones((4, 5)))
    assert ds["rA"].dims == ("Y", "X") and float(np.asarray(ds["rA"]).sum()) == 20.0
    ds["Vol"] = ds["drF"] * ds["rA"]
    assert ds["Vol"].dims == ("Z", "Y", "X")
    ds["face_index"] = xr.DataArray(np.arange(5), dims="X")
    assert ds["face_index"].dims == ("X",)
