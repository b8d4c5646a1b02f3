"""Reference-produced golden vectors (tools/ref_vectors, run once on a machine with cargo): proof bytes, commitments and
verdicts of the REAL reference under fixed blindings / RNG bytes. When tests/golden/reference_*.json exist the oracle must
reproduce them bit for bit — that pins every transcript label, the draw order and the R1CSProof layout (SURVEY.md §8c risks
R1, R2). Without them the oracle stays pinned to third-party primitives only: parity unpinned at the proof-byte level."""
import glob
import json
import os

import pytest

import orc

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_*.json")))


def load_vectors():
    out = []
    for path in GOLDEN:
        for v in json.load(open(path))["vectors"]:
            out.append(v)
    return out


def as_bid(v):
    h = bytes.fromhex
    return dict(d=h(v["d"]), k=h(v["k"]), y=h(v["y"]), y_inv=h(v["y_inv"]), q=h(v["q"]), z_img=h(v["z_img"]), seed=h(v["seed"]),
                pub_list=b"".join(h(x) for x in v["pub_list"]), L=v["L"], toggle=v["toggle"], blindings=b"".join(h(x) for x in v["blindings"]),
                rng_seed=h(v["rng_seed"]))


@pytest.mark.skipif(not GOLDEN, reason="parity unpinned: no reference-produced vectors under tests/golden/reference_*.json (tools/ref_vectors emits them)")
def test_oracle_reproduces_reference_vectors():
    for v in load_vectors():
        bid = as_bid(v)
        versioned = 1 if v["proof_len"] == len(bytes.fromhex(v["proof"])) and v["proof_len"] % 32 == 1 else 0   # R1: which blob layout the reference emits
        rc, proof, comm, tc = orc.blindbid_prove(bid, bid["blindings"], bid["rng_seed"], versioned=versioned)
        assert rc == 0
        assert comm.hex() == "".join(v["commitments"]) and tc.hex() == "".join(v["t_c"]), "commitments differ from the reference"
        assert proof.hex() == v["proof"], "proof bytes differ from the reference (label, draw order or layout)"
        got = orc.blindbid_verify(bytes.fromhex(v["proof"]), comm, tc, bid["q"], bid["z_img"], bid["seed"], bid["pub_list"], bytes(32), versioned=versioned)
        assert (got == 0) == v["verdict"]


def test_kit_is_present_and_states_its_recipe():
    kit = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "ref_vectors")
    assert os.path.exists(os.path.join(kit, "Cargo.toml")) and os.path.exists(os.path.join(kit, "src", "main.rs"))
    assert "4a05305095abe2643122184deadf288fd85617a3" in open(os.path.join(kit, "Cargo.toml")).read()
