"""Pins oracle/philox.py to the Random123 known-answer vectors for
philox4x32-10 and checks the draw spec's basic statistics."""
import numpy as np

from oracle import philox as P


def _kat(ctr, key):
    out = P.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
    return [int(x) for x in out]


def test_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    assert _kat([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert _kat([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert _kat([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_rademacher_and_gaussian_stats():
    r = P.rademacher(7, 5, 20000)
    assert r.shape == (5, 20000) and r.dtype == np.float32
    assert set(np.unique(r)) == {-1.0, 1.0}
    assert abs(r.mean()) < 0.02
    g = P.gaussian(7, 5, 20000)
    assert g.shape == (5, 20000) and np.isfinite(g).all()
    assert abs(g.mean()) < 0.02 and abs(g.std() - 1.0) < 0.02
    c = np.corrcoef(g)
    assert np.abs(c - np.eye(5)).max() < 0.03


def test_sharding_invariance_and_streams():
    full = P.gaussian(3, 6, 64)
    a = P.gaussian(3, 6, 32, offset=0)
    b = P.gaussian(3, 6, 32, offset=32)
    assert np.array_equal(full, np.concatenate([a, b], axis=1))
    assert not np.array_equal(P.gaussian(3, 6, 8), P.gaussian(3, 6, 8, stream=P.STREAM_BASE))
    assert not np.array_equal(P.gaussian(3, 6, 8), P.gaussian(4, 6, 8))
    u = P.uniform_pm(11, 0.1)
    assert -0.1 <= u <= 0.1
