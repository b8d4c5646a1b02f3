"""GPU parity, aggregated range proofs (BASELINE config 5): proof bytes and commitments identical to
the oracle's under the same blindings / rng stream, identical accept / reject decisions, at a small size and at the
config-5 shape (m = 64 parties x 64 bits: 4096-element IPP, 1056-byte proof)."""
import ctypes
import hashlib
import os

import pytest

import orc
from orc import L_ORDER, from_le, le

pytestmark = pytest.mark.gpu
RNG = bytes(range(32))


def blindings(m, tag=b"rp-bl"):
    return b"".join(le(from_le(hashlib.shake_256(tag + bytes([i])).digest(64)) % L_ORDER) for i in range(m))


def oracle_prove(values, nbits, bl, seed):
    lib = orc.lib()
    m = len(values)
    vals = (ctypes.c_uint64 * m)(*values)
    proof, plen, V = ctypes.create_string_buffer(4096), ctypes.c_size_t(4096), ctypes.create_string_buffer(32 * m)
    rc = lib.orc_rangeproof_prove(vals, bl, ctypes.c_size_t(m), ctypes.c_size_t(nbits), seed, proof, ctypes.byref(plen), V)
    return rc, proof.raw[:plen.value], V.raw


def oracle_verify(proof, V, nbits, seed):
    return orc.lib().orc_rangeproof_verify(proof, ctypes.c_size_t(len(proof)), V, ctypes.c_size_t(len(V) // 32), ctypes.c_size_t(nbits), seed, 4)


def backend(nbits, parties):
    import bbp_loader
    return bbp_loader.load().Backend(device=0, gens_capacity=nbits, party_capacity=parties)


@pytest.mark.parametrize("replay", ["host", "device"])
def test_rangeproof_small_matches_oracle(replay, monkeypatch):
    # the verifier's transcript replay runs on the host for small batches and one warp per proof on the device for large
    # ones (BBP_DEVICE_TRANSCRIPT_MIN_BATCH is the crossover): force each, same verdicts
    monkeypatch.setenv("BBP_DEVICE_TRANSCRIPT_MIN_BATCH", "1000000" if replay == "host" else "1")
    be = backend(8, 4)
    seed = b"\x07" * 32
    for values in ([0, 255, 17, 128], [3, 200], [77]):
        m = len(values)
        bl = blindings(m)
        rc, proof, V = oracle_prove(values, 8, bl, seed)
        grc, gproof, gV = be.rangeproof_prove(values, bl, 8, seed)
        assert rc == 0 and grc == 0
        assert gV == V
        assert gproof == proof
        assert be.rangeproof_verify(proof, V, 8, RNG) == 0 and oracle_verify(gproof, gV, 8, RNG) == 0
        # mutations: same verdict class as the oracle
        for pos in (0, 40, 70, 100, 130, 170, 200, 230, len(proof) - 40, len(proof) - 1):
            bad = bytearray(proof)
            bad[pos] ^= 2
            assert be.rangeproof_verify(bytes(bad), V, 8, RNG) == oracle_verify(bytes(bad), V, 8, RNG) != 0, pos
        assert be.rangeproof_verify(proof[:-32], V, 8, RNG) == oracle_verify(proof[:-32], V, 8, RNG) != 0
        badV = bytearray(V)
        badV[0:32] = V[32:64] if m > 1 else bytes(32)
        assert be.rangeproof_verify(proof, bytes(badV), 8, RNG) == oracle_verify(proof, bytes(badV), 8, RNG) != 0
    # a value that does not fit: proof is produced, must not verify (both sides)
    bl = blindings(2)
    rc, proof, V = oracle_prove([256, 1], 8, bl, seed)
    grc, gproof, gV = be.rangeproof_prove([256, 1], bl, 8, seed)
    assert (grc, gproof, gV) == (rc, proof, V)
    assert be.rangeproof_verify(gproof, gV, 8, RNG) == oracle_verify(proof, V, 8, RNG) == -3
    # generator capacity: 8 parties do not fit a (8, 4) context
    grc, _, _ = be.rangeproof_prove(list(range(8)), blindings(8), 8, seed)
    assert grc == -1
    be.close()


@pytest.mark.parametrize("hybrid,device_rng", [("0", "1000000"), ("2", "1000000"), ("2", "1")])
def test_rangeproof_config5_shape(hybrid, device_rng, monkeypatch):
    """m = 64, n = 64: 4096-element IPP, 12 rounds, 1056-byte proof, byte-identical to the oracle; batch of 3; with the
    plain and the hybrid (materialised bases) inner-product argument, and with the SHAKE256 draw stream squeezed on the
    host or on the device"""
    monkeypatch.setenv("BBP_IPP_HYBRID", hybrid)
    monkeypatch.setenv("BBP_DEVICE_RNG_MIN_BATCH", device_rng)
    be = backend(64, 64)
    seeds = [hashlib.sha256(b"rp%d" % i).digest() for i in range(3)]
    vals = [[from_le(hashlib.shake_256(b"rp-v" + bytes([i, k])).digest(8)) for i in range(64)] for k in range(3)]
    bls = [blindings(64, b"rp-bl%d" % k) for k in range(3)]
    st, proofs, Vs = be.rangeproof_prove_batch(vals, b"".join(bls), 64, 64, b"".join(seeds))
    assert st == [0, 0, 0]
    rc, proof, V = oracle_prove(vals[0], 64, bls[0], seeds[0])
    assert rc == 0 and len(proof) == 1056
    assert proofs[0] == proof and Vs[0] == V
    assert oracle_verify(proofs[2], Vs[2], 64, RNG) == 0
    bad = [proofs[0], proofs[1][:200] + bytes([proofs[1][200] ^ 1]) + proofs[1][201:], proofs[2]]
    want_bad = [0, oracle_verify(bad[1], Vs[1], 64, RNG), 0]
    assert want_bad[1] != 0
    # batches are first checked as ONE random linear combination (accept all if it is the identity), then per request if
    # that fails; BBP_RP_COMBINED=0 goes straight to the per-request pass: same verdicts either way
    for combined, replay_min in (("1", "1000000"), ("0", "1000000"), ("1", "1"), ("0", "1")):
        monkeypatch.setenv("BBP_RP_COMBINED", combined)
        monkeypatch.setenv("BBP_DEVICE_TRANSCRIPT_MIN_BATCH", replay_min)   # host / device transcript replay
        assert be.rangeproof_verify_batch(proofs, Vs, 64, 64, RNG * 3) == [0, 0, 0]
        assert be.rangeproof_verify_batch(bad, Vs, 64, 64, RNG * 3) == want_bad
        # a commitment swapped between two requests breaks both
        swapped = [Vs[1], Vs[0], Vs[2]]
        assert be.rangeproof_verify_batch(proofs, swapped, 64, 64, RNG * 3) == [-3, -3, 0]
    be.close()


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_ipp_partition_gives_identical_proofs(shards, monkeypatch):
    """BASELINE config 5's sharded inner-product argument, the partition emulated on one GPU (bbp_set_ipp_shard(emulate)):
    generator columns i = g (mod G) per shard, per-round sum of the shards' partial L / R points -> the proof bytes of the
    unsharded prover and of the oracle (m = 64 x 64 bits: 4096-element IPP, 12 rounds)."""
    be = backend(64, 64)
    seeds = [hashlib.sha256(b"sh%d" % i).digest() for i in range(2)]
    vals = [[from_le(hashlib.shake_256(b"sh-v" + bytes([i, k])).digest(8)) for i in range(64)] for k in range(2)]
    bls = [blindings(64, b"sh-bl%d" % k) for k in range(2)]
    st, plain, Vs = be.rangeproof_prove_batch(vals, b"".join(bls), 64, 64, b"".join(seeds))
    assert st == [0, 0]
    l0 = be.launch_count()
    be.set_ipp_shard(0, 1, None, emulate=shards)
    st, sharded, Vs2 = be.rangeproof_prove_batch(vals, b"".join(bls), 64, 64, b"".join(seeds))
    be.set_ipp_shard(0, 1, None, emulate=0)
    assert st == [0, 0] and sharded == plain and Vs2 == Vs
    assert be.launch_count() - l0 > 12 * shards            # one MSM per shard and round actually ran
    rc, proof, V = oracle_prove(vals[1], 64, bls[1], seeds[1])
    assert rc == 0 and sharded[1] == proof
    be.close()


def test_sharded_ipp_two_ranks_nccl():
    """the same over NCCL: two processes, one GPU each, identical inputs, per-round all-gather of the partial sums; every rank
    returns the single-GPU proof. Needs two devices (skipped on a one-GPU box)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sharded_ipp_worker.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", script], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("SHARDED-IPP-OK") == 2
