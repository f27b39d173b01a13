"""CPU: the host side of the result pipeline.  ptfnn_savetxt must write the bytes np.savetxt
writes for the reference's formats (R:454-481, R:855-860) and ptfnn_loadtxt must return what np.loadtxt
returns for those files (R:794-831) -- NumPy itself is the reference implementation here."""
import os

import numpy as np
import pytest

from ptnn_b200 import capi

FORMATS = ['%.18e', '%1.8f', '%1.2f', '%1.4f', '%1.5f']          # every fmt the reference passes to np.savetxt


def _values(n=20000):
    rs = np.random.RandomState(0)
    x = np.concatenate([rs.randn(n), rs.randn(n) * 10.0 ** rs.randint(-12, 12, n),
                        np.frombuffer(rs.bytes(8 * n), dtype=np.float64),
                        [0.0, -0.0, 0.5, 1.5, 2.5, 0.125, 0.005, 0.015, 1.005, 2.675, -100.0, 1e22, 5e-324, 1.7976931348623157e308]])
    return x[np.isfinite(x)]


@pytest.mark.parametrize("fmt", FORMATS + ['%10.3g', '%+.6e'])
def test_savetxt_writes_numpy_bytes_and_loadtxt_reads_them_back(tmp_path, fmt):
    x = _values()
    if 'f' in fmt:
        x = x[np.abs(x) < 1e15]
    for k, X in enumerate((x, x[:9000].reshape(-1, 3), x[:31].reshape(1, 31), x[:1], np.ones((4, 2)))):
        a, b = str(tmp_path / ("np%d.txt" % k)), str(tmp_path / ("our%d.txt" % k))
        np.savetxt(a, X, fmt=fmt)
        capi.savetxt(b, X, fmt=fmt)
        assert open(a, "rb").read() == open(b, "rb").read(), (fmt, k)
        want, got = np.loadtxt(a), capi.loadtxt(b)
        assert want.shape == got.shape and np.array_equal(want, got), (fmt, k)


def test_non_finite_values_and_default_format(tmp_path):
    X = np.array([[np.inf, -np.inf, np.nan], [1.0, -2.0, 3.0]])
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    np.savetxt(a, X)
    capi.savetxt(b, X)                                           # default fmt = np.savetxt's '%.18e' (R:455)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert np.array_equal(np.loadtxt(a), capi.loadtxt(b), equal_nan=True)


def test_errors(tmp_path):
    with pytest.raises(OSError):
        capi.loadtxt(str(tmp_path / "missing.txt"))
    with pytest.raises(OSError):
        capi.savetxt(str(tmp_path / "no_such_dir" / "x.txt"), np.ones(3))
    with pytest.raises(OSError):
        capi.savetxt(str(tmp_path / "x.txt"), np.ones(3), fmt='%d %s')      # one float conversion only
    with pytest.raises(ValueError):
        capi.savetxt(str(tmp_path / "x.txt"), np.ones((2, 2, 2)))
    p = tmp_path / "ragged.txt"
    p.write_text("1 2 3\n4 5\n")
    with pytest.raises(OSError):
        capi.loadtxt(str(p))
    p = tmp_path / "blank.txt"
    p.write_text("\n1.5 2\n\n3 4\n")
    assert np.array_equal(capi.loadtxt(str(p)), [[1.5, 2.0], [3.0, 4.0]])
    assert not os.path.exists(str(tmp_path / "no_such_dir"))
