"""
HDF5 layout of the result containers, checked without h5py (absent from the image)
through a dict-backed stand-in for `h5py.Group`:

  * the key tree, dataset shapes and dtypes written by `CorrFunc.to_hdf` are the reference's
    (`src/yaw/correlation/paircounts.py:226-232, 394-408`, `corrfunc.py:174-181, 323-325`);
  * `to_hdf -> from_hdf` round-trips bit for bit;
  * with the reference checkout present, the UNMODIFIED reference writes the same counts into the same kind of
    fake group and the two trees are compared key by key, value by value.
"""

import numpy as np
import pytest
from numpy.testing import assert_array_equal

import golden_cases
import golden_io
import refshim
from fake_engine import OracleEngine


class FakeDataset:
    def __init__(self, data, **kwargs):
        if isinstance(data, str):  # h5py stores text as bytes and returns bytes on read
            data = data.encode("utf-8")
        self.data = np.asarray(data)
        self.kwargs = kwargs  # compression options

    def __getitem__(self, key):
        out = self.data[key]
        return out.item() if (key == () and self.data.ndim == 0) else out

    @property
    def shape(self):
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype


class FakeGroup(dict):
    """the subset of `h5py.Group` the containers use: create_group / create_dataset / __getitem__ / __contains__"""

    def create_group(self, name):
        assert name not in self, name
        self[name] = FakeGroup()
        return self[name]

    def create_dataset(self, name, data=None, **kwargs):
        assert name not in self, name
        self[name] = FakeDataset(data, **kwargs)
        return self[name]

    def tree(self, prefix=""):
        out = {}
        for key, val in self.items():
            if isinstance(val, FakeGroup):
                out.update(val.tree(f"{prefix}{key}/"))
            else:
                out[f"{prefix}{key}"] = val
        return out


@pytest.fixture(scope="module")
def corrfunc():
    g = golden_io.load("cross_weighted_multiscale")
    return g, golden_cases.run_cross(g, OracleEngine())


def test_corrfunc_hdf_key_tree_and_shapes(corrfunc):
    g, corrs = corrfunc
    cf = corrs[0]
    root = FakeGroup()
    cf.to_hdf(root)
    tree = root.tree()
    n_bins, n_patch = cf.dd.counts.num_bins, cf.dd.counts.num_patches
    # reference layout: /{data_data,data_random,random_data,random_random}/{counts/..., sum_weights/...} + version, kind
    assert tree["kind"][()].decode() == "CorrFunc" and "version" in tree
    for name in ("data_data", "data_random", "random_data", "random_random"):
        assert f"{name}/version" in tree
        for key in ("binning/closed", "binning/edges", "auto", "num_patches", "patch_pairs", "binned_counts", "version"):
            assert f"{name}/counts/{key}" in tree, (name, key)
        for key in ("binning/closed", "binning/edges", "auto", "sum_weights1", "sum_weights2", "version"):
            assert f"{name}/sum_weights/{key}" in tree, (name, key)
        pairs, binned = tree[f"{name}/counts/patch_pairs"], tree[f"{name}/counts/binned_counts"]
        assert pairs.shape[1] == 2 and binned.shape == (pairs.shape[0], n_bins)
        assert int(tree[f"{name}/counts/num_patches"][()]) == n_patch
        assert tree[f"{name}/counts/binning/edges"].shape == (n_bins + 1,)
        assert tree[f"{name}/sum_weights/sum_weights1"].shape == (n_bins, n_patch)
        assert tree[f"{name}/sum_weights/sum_weights2"].shape == (n_bins, n_patch)
        # only patch pairs with any non-zero bin are stored (paircounts.py:401-408)
        assert np.all(np.any(binned.data != 0, axis=1))
        # the large arrays carry the reference's compression options (utils/misc.py:36)
        for key in ("patch_pairs", "binned_counts"):
            assert tree[f"{name}/counts/{key}"].kwargs == dict(fletcher32=True, compression="gzip", shuffle=True)
    assert set(k.split("/")[0] for k in tree) == {"version", "kind", "data_data", "data_random", "random_data", "random_random"}


def test_corrfunc_hdf_round_trip(corrfunc):
    import yet_another_wizz_b200 as yb

    _, corrs = corrfunc
    for cf in corrs:
        root = FakeGroup()
        cf.to_hdf(root)
        back = yb.CorrFunc.from_hdf(root)
        assert back == cf
        for kind in ("dd", "dr", "rd", "rr"):
            assert_array_equal(getattr(back, kind).counts.counts, getattr(cf, kind).counts.counts)
            assert_array_equal(getattr(back, kind).sum_weights.sum_weights1, getattr(cf, kind).sum_weights.sum_weights1)
            assert getattr(back, kind).counts.auto == getattr(cf, kind).counts.auto


def test_autocorrelation_hdf_round_trip():
    import yet_another_wizz_b200 as yb

    g = golden_io.load("auto_unweighted")
    cf = golden_cases.run_auto(g, OracleEngine())[0]
    root = FakeGroup()
    cf.to_hdf(root)
    assert "random_data" not in root  # autocorrelate returns CorrFunc(dd, dr, None, rr)
    assert bool(root["data_data"]["counts"]["auto"][()]) is True
    back = yb.CorrFunc.from_hdf(root)
    assert back == cf and back.rd is None


@pytest.mark.reference
@pytest.mark.skipif(not refshim.reference_available(), reason="reference checkout not present")
def test_hdf_tree_equals_reference(corrfunc):
    """the unmodified reference, given the same counts (`CorrFunc.to_reference`), writes the same tree"""
    refshim.import_reference()
    _, corrs = corrfunc
    cf = corrs[0]
    mine, ref = FakeGroup(), FakeGroup()
    cf.to_hdf(mine)
    cf.to_reference().to_hdf(ref)
    a, b = mine.tree(), ref.tree()
    assert set(a) == set(b)
    for key in a:
        va, vb = a[key].data, b[key].data
        if key.endswith("version"):
            continue  # the version tag names the writing package
        assert va.shape == vb.shape and va.dtype.kind == vb.dtype.kind, key
        if va.dtype.kind in "US":
            assert va.item() == vb.item(), key
        else:
            assert_array_equal(va, vb, err_msg=key)
        assert a[key].kwargs == b[key].kwargs, key
