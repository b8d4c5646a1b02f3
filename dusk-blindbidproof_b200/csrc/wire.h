// Outer (process) boundary of the reference: one TLV frame per request over a Unix socket, payload byte 0 = opcode
// (src/futures/main.rs:64-110). This file holds the byte-level codec — request / reply (de)serialisation of
// src/blindbid/proof.rs:97-184, src/blindbid/verify.rs:91-128 and src/blindbid/bid.rs:15-29 — without any socket or GPU
// code, so that it can be driven from tests through the C ABI (bbp_wire_*) and from the server shell (server/).
//
// FRAMING IS UNPINNED. The reference delegates framing to the crate dusk-tlv 1.0.1 @ 5be856b (Cargo.lock:183-185), whose
// source is not under /root/reference and which no fixture in the reference pins. What is implemented is the recollected
// form: an item is one tag byte giving the width of the length field (1, 2, 4 or 8), the length in
// little-endian, then the payload; a list is one item whose payload is the concatenation of its elements' items; the serde
// path writes a Scalar as one 32-byte item and a u64 as one 8-byte little-endian item. Everything framing-specific sits in
// tlv_put / tlv_get below; the request / reply layouts above them follow the reference's code line by line.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace bbp {
namespace wire {

typedef std::vector<uint8_t> bytes;

// ---- dusk-tlv framing (unpinned, see header) ---------------------------------------------------------------------------
inline void tlv_put(bytes &out, const uint8_t *payload, size_t len) {
    int w = len < (1ull << 8) ? 1 : len < (1ull << 16) ? 2 : len < (1ull << 32) ? 4 : 8;
    out.push_back((uint8_t)w);
    for (int i = 0; i < w; i++) out.push_back((uint8_t)((uint64_t)len >> (8 * i)));
    out.insert(out.end(), payload, payload + len);
}
inline void tlv_put(bytes &out, const bytes &payload) { tlv_put(out, payload.data(), payload.size()); }
// header of the item at in[0..len): false on a bad tag or a truncated header; does NOT require the payload to be present
inline bool tlv_header(const uint8_t *in, size_t len, size_t *hdr, uint64_t *payload_len) {
    if (len < 1) return false;
    int w = in[0];
    if (w != 1 && w != 2 && w != 4 && w != 8) return false;
    if (len < 1 + (size_t)w) return false;
    uint64_t l = 0;
    for (int i = 0; i < w; i++) l |= (uint64_t)in[1 + i] << (8 * i);
    *hdr = 1 + (size_t)w;
    *payload_len = l;
    return true;
}
struct reader {
    const uint8_t *p;
    size_t n;
    bool ok = true;
    reader(const uint8_t *data, size_t len) : p(data), n(len) {}
    bool at_end() const { return n == 0; }
    // next item: pointer into the buffer + length; ok = false on malformed / truncated input
    bool next(const uint8_t **payload, size_t *len) {
        size_t hdr;
        uint64_t l;
        if (!ok || !tlv_header(p, n, &hdr, &l) || l > n - hdr) { ok = false; return false; }
        *payload = p + hdr;
        *len = (size_t)l;
        p += hdr + (size_t)l;
        n -= hdr + (size_t)l;
        return true;
    }
    // TlvReader::read_list::<Vec<u8>>(): one item holding the elements' items
    bool list(std::vector<std::pair<const uint8_t *, size_t>> &out) {
        const uint8_t *body;
        size_t len;
        if (!next(&body, &len)) return false;
        reader in(body, len);
        while (!in.at_end()) {
            const uint8_t *e;
            size_t el;
            if (!in.next(&e, &el)) { ok = false; return false; }
            out.push_back({e, el});
        }
        return true;
    }
    // serde path: Scalar = one 32-byte item (canonicity is checked by the caller), u64 = one 8-byte LE item
    bool scalar32(const uint8_t **s) {
        size_t len;
        if (!next(s, &len) || len != 32) { ok = false; return false; }
        return true;
    }
    bool u64(uint64_t *x) {
        const uint8_t *b;
        size_t len;
        if (!next(&b, &len) || len != 8) { ok = false; return false; }
        *x = 0;
        for (int i = 0; i < 8; i++) *x |= (uint64_t)b[i] << (8 * i);
        return true;
    }
};
inline void put_list32(bytes &out, const uint8_t *items, size_t count) {
    bytes body;
    for (size_t i = 0; i < count; i++) tlv_put(body, items + 32 * i, 32);
    tlv_put(out, body);
}

// ---- requests ----------------------------------------------------------------------------------------------------------
enum : int { OP_PROVE = 1, OP_VERIFY = 2 };

struct request {
    int opcode = 0;
    // prove (src/blindbid/proof.rs:97-114): d, k, y, y_inv, q, z_img, seed, the bid list, toggle
    uint8_t scalars[7][32];
    bytes pub_list;          // L x 32
    uint64_t toggle = 0;
    // verify (src/blindbid/verify.rs:91-128): the proof blob's three parts, score / z_img / seed (scalars[0..3)), the list
    bytes proof, commitments, t_c;
};

// Proof::try_from (proof.rs:145-184): TLV(R1CSProof bytes) | list(commitments) | list(t_c), 32-byte elements only
inline bool parse_proof_blob(const uint8_t *data, size_t len, bytes &proof, bytes &commitments, bytes &t_c) {
    reader r(data, len);
    const uint8_t *p;
    size_t pl;
    if (!r.next(&p, &pl)) return false;
    proof.assign(p, p + pl);
    for (bytes *dst : {&commitments, &t_c}) {
        std::vector<std::pair<const uint8_t *, size_t>> items;
        if (!r.list(items)) return false;
        for (auto &it : items) {
            if (it.second != 32) return false;   // "Compressed Ristrettos can only be created from 32 bytes slices"
            dst->insert(dst->end(), it.first, it.first + 32);
        }
    }
    return true;
}

// payload of the request frame (opcode byte + body) -> request. Returns the opcode, 0 for an unknown opcode (the
// reference writes nothing: "Undefined operation code"), -1 for a body the reference's readers would reject.
inline int parse_request(const uint8_t *payload, size_t len, request &R) {
    if (len < 1) return -1;
    R.opcode = payload[0];
    reader r(payload + 1, len - 1);
    if (R.opcode == OP_PROVE) {
        for (int i = 0; i < 7; i++) {
            const uint8_t *s;
            if (!r.scalar32(&s)) return -1;
            memcpy(R.scalars[i], s, 32);
        }
        std::vector<std::pair<const uint8_t *, size_t>> items;
        if (!r.list(items)) return -1;
        for (auto &it : items) {
            if (it.second != 32) return -1;      // Bid::from panics on any other length (bid.rs:23-25): no reply either way
            R.pub_list.insert(R.pub_list.end(), it.first, it.first + 32);
        }
        if (!r.u64(&R.toggle)) return -1;
        return OP_PROVE;
    }
    if (R.opcode == OP_VERIFY) {
        const uint8_t *blob;
        size_t bl;
        if (!r.next(&blob, &bl)) return -1;
        if (!parse_proof_blob(blob, bl, R.proof, R.commitments, R.t_c)) return -1;
        for (int i = 0; i < 3; i++) {
            const uint8_t *s;
            if (!r.scalar32(&s)) return -1;
            memcpy(R.scalars[i], s, 32);
        }
        std::vector<std::pair<const uint8_t *, size_t>> items;
        if (!r.list(items)) return -1;
        for (auto &it : items) {
            if (it.second != 32) return -1;      // "Scalars Ristrettos can only be created from 32 bytes slices"
            R.pub_list.insert(R.pub_list.end(), it.first, it.first + 32);
        }
        return OP_VERIFY;
    }
    return 0;
}

// ---- replies and client-side encoders -------------------------------------------------------------------------------------
// Proof::try_into (proof.rs:118-143), wrapped in the reply frame (futures/main.rs:87-92)
inline bytes proof_blob(const uint8_t *proof, size_t proof_len, const uint8_t *commitments, size_t nc, const uint8_t *t_c, size_t nt) {
    bytes body;
    tlv_put(body, proof, proof_len);
    put_list32(body, commitments, nc);
    put_list32(body, t_c, nt);
    return body;
}
inline bytes prove_reply(const uint8_t *proof, size_t proof_len, const uint8_t *commitments, size_t nc, const uint8_t *t_c, size_t nt) {
    bytes out;
    tlv_put(out, proof_blob(proof, proof_len, commitments, nc, t_c, nt));
    return out;
}
// one byte: 0x01 accept, 0x00 reject or any parse / format error (futures/main.rs:95-101)
inline bytes verify_reply(bool accept) {
    bytes out;
    uint8_t b = accept ? 1 : 0;
    tlv_put(out, &b, 1);
    return out;
}
// what the Go client sends (the frames the parsers above accept)
inline bytes prove_request(const uint8_t scalars[7][32], const uint8_t *pub_list, size_t L, uint64_t toggle) {
    bytes body;
    body.push_back((uint8_t)OP_PROVE);
    for (int i = 0; i < 7; i++) tlv_put(body, scalars[i], 32);
    put_list32(body, pub_list, L);
    uint8_t t[8];
    for (int i = 0; i < 8; i++) t[i] = (uint8_t)(toggle >> (8 * i));
    tlv_put(body, t, 8);
    bytes out;
    tlv_put(out, body);
    return out;
}
inline bytes verify_request(const uint8_t *blob, size_t blob_len, const uint8_t score[32], const uint8_t z_img[32], const uint8_t seed[32],
                            const uint8_t *pub_list, size_t L) {
    bytes body;
    body.push_back((uint8_t)OP_VERIFY);
    tlv_put(body, blob, blob_len);
    tlv_put(body, score, 32);
    tlv_put(body, z_img, 32);
    tlv_put(body, seed, 32);
    put_list32(body, pub_list, L);
    bytes out;
    tlv_put(out, body);
    return out;
}

}  // namespace wire
}  // namespace bbp
