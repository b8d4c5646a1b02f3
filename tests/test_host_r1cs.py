"""CPU: the flattened-circuit front end of the generic bulletproofs surface (bbp_cs_shape = the validation + template build
bbp_r1cs_prove / _verify run first), and the transcript object, without a GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbp_loader  # noqa: E402
from orc import L_ORDER, le  # noqa: E402
from r1cs_util import LC, Recorder, example_circuit  # noqa: E402

capi = bbp_loader.load().capi


def test_shape_of_recorded_circuits():
    for n_extra in (0, 3, 40):
        cs = example_circuit(5, n_extra)
        flat = cs.flatten()
        rc, shape = capi.cs_shape(flat)
        n1 = 2 + n_extra
        assert rc == 0 and shape[:3] == [n1, 2 * n1 + 3, 3]
        assert shape[3] == 1 << (n1 - 1).bit_length()
        assert shape[4] > 1                      # non +-1 coefficients and constants landed in the table
    # a circuit whose coefficients are all +-1 needs no table beyond the entry for one
    cs = Recorder([3, 4])
    _, _, o = cs.multiply(cs.committed(0), cs.committed(1))
    cs.constrain(o - cs.committed(0) - cs.committed(0) - cs.committed(0) - cs.committed(0))
    assert capi.cs_shape(cs.flatten()) == (0, [1, 3, 2, 1, 1])


def test_malformed_circuits_are_refused():
    cs = example_circuit(6, 2)
    flat = cs.flatten()
    bad = dict(flat); bad["term_var"] = [(1 << 28) | 99] + flat["term_var"][1:]          # multiplier index out of range
    assert capi.cs_shape(bad)[0] == capi.BBP_ERR_FORMAT
    bad = dict(flat); bad["term_var"] = [(0 << 28) | 3] + flat["term_var"][1:]           # commitment index out of range
    assert capi.cs_shape(bad)[0] == capi.BBP_ERR_FORMAT
    bad = dict(flat); bad["term_var"] = [(7 << 28)] + flat["term_var"][1:]               # unknown variable kind
    assert capi.cs_shape(bad)[0] == capi.BBP_ERR_FORMAT
    bad = dict(flat); bad["term_coeff"] = le(L_ORDER) + flat["term_coeff"][32:]          # non-canonical coefficient
    assert capi.cs_shape(bad)[0] == capi.BBP_ERR_FORMAT
    bad = dict(flat); bad["con_ptr"] = [0, 5, 3] + flat["con_ptr"][3:]                   # decreasing row pointer
    assert capi.cs_shape(bad)[0] == capi.BBP_ERR_FORMAT
    bad = dict(flat); bad["n_mul"] = 0
    assert capi.cs_shape(bad)[0] == capi.BBP_ERR_INPUT


def test_transcript_object_merlin_vector_on_cpu():
    t = capi.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_cs_shape_on_random_garbage():
    """random term lists are either accepted with a consistent shape or refused — never a crash"""
    import random
    rnd = random.Random(99)
    ok = refused = 0
    for _ in range(300):
        n_mul, m, q = rnd.randrange(1, 9), rnd.randrange(0, 4), rnd.randrange(0, 12)
        con_ptr = [0]
        for _ in range(q):
            con_ptr.append(con_ptr[-1] + rnd.randrange(0, 5))
        nt = con_ptr[-1]
        term_var = [(rnd.randrange(0, 6) << 28) | rnd.randrange(0, 10) for _ in range(nt)]
        coeff = b"".join(le(rnd.choice([0, 1, L_ORDER - 1, rnd.getrandbits(252) % L_ORDER, L_ORDER + rnd.randrange(3)]) % (1 << 256)) for _ in range(nt))
        rc, shape = capi.cs_shape(dict(n_mul=n_mul, m=m, con_ptr=con_ptr, term_var=term_var, term_coeff=coeff or b"\0"))
        if rc == 0:
            ok += 1
            assert shape[:3] == [n_mul, q, m]
        else:
            refused += 1
            assert rc == capi.BBP_ERR_FORMAT
    assert ok > 5 and refused > 5
