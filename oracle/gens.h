// ORACLE — test infrastructure only (see fe.h header).
// bulletproofs@4a05305 generators, as used by
// generate_cs_transcript() at src/blindbid/mod.rs:34-40: PedersenGens::default() and
// BulletproofGens::new(gens_capacity, party_capacity).
#pragma once
#include "ge.h"
#include "hash.h"
#include <vector>

namespace orc {

struct pedersen_gens {
    ge B, B_blinding;
    pedersen_gens() {
        B = ge_basepoint();
        uint8_t comp[32], h[64];
        ge_compress(comp, B);
        sha3_512(h, comp, 32);
        B_blinding = ge_from_uniform_bytes(h);   // RistrettoPoint::hash_from_bytes::<Sha3_512>
    }
    ge commit(const sc &v, const sc &blinding) const {
        return ge_add(ge_scalarmul(v, B), ge_scalarmul(blinding, B_blinding));
    }
};

// GeneratorsChain: SHAKE256("GeneratorsChain" || label) read in 64-byte blocks
static inline void generators_chain(std::vector<ge> &out, char which, uint32_t party, size_t count) {
    shake256 sh;
    sh.absorb((const uint8_t *)"GeneratorsChain", 15);
    uint8_t label[5] = {(uint8_t)which, (uint8_t)party, (uint8_t)(party >> 8), (uint8_t)(party >> 16), (uint8_t)(party >> 24)};
    sh.absorb(label, 5);
    out.resize(count);
    for (size_t i = 0; i < count; i++) {
        uint8_t blk[64];
        sh.squeeze(blk, 64);
        out[i] = ge_from_uniform_bytes(blk);
    }
}

struct bulletproof_gens {
    size_t gens_capacity, party_capacity;
    std::vector<std::vector<ge>> G, H;   // [party][i]
    bulletproof_gens(size_t gens_cap, size_t party_cap) : gens_capacity(gens_cap), party_capacity(party_cap), G(party_cap), H(party_cap) {
        for (size_t j = 0; j < party_cap; j++) {
            generators_chain(G[j], 'G', (uint32_t)j, gens_cap);
            generators_chain(H[j], 'H', (uint32_t)j, gens_cap);
        }
    }
};

}  // namespace orc
