"""CPU checks of the product's host-side circuit logic (no GPU): MiMC constants and hash against the golden values and
the oracle, the circuit shape n1 = 1442 + 3L, q = 2887 + 9L, m = 4 + L."""
import ctypes
import hashlib
import json
import os

import pytest

import bbp_loader
import orc

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ristretto_libsodium.json")))
capi = bbp_loader.load().capi


def test_mimc_constants_golden():
    c = capi.mimc_constants()
    assert c[:32].hex() == GOLD["mimc_constants_first"]
    assert c[89 * 32:].hex() == GOLD["mimc_constants_last"]
    assert hashlib.sha512(c).hexdigest() == GOLD["mimc_constants_sha512"]


def test_mimc_hash_matches_oracle():
    for i in range(20):
        a = orc.random_scalars(100 + i, 1)
        b = orc.random_scalars(200 + i, 1)
        assert capi.mimc_hash(a, b) == orc.mimc_hash(a, b)
    assert capi.mimc_hash(bytes(32), bytes(32)) == orc.mimc_hash(bytes(32), bytes(32))


@pytest.mark.parametrize("L", [1, 2, 8, 64, 202, 203])
def test_circuit_shape(L):
    n1, q, m = capi.circuit_shape(4, L)
    assert (n1, q, m) == (1442 + 3 * L, 2887 + 9 * L, 4 + L)
    out = (ctypes.c_size_t * 3)()
    orc.lib().orc_blindbid_shape(ctypes.c_size_t(L), out)
    assert (out[0], out[1], out[2]) == (n1, q, m)
