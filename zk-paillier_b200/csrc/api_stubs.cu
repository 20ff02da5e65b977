// Temporary: entry points not implemented yet report ZKP_E_STATE.
#include "ctx.h"
using namespace zkp;
extern "C" {
#define NI(c) return fail(c, ZKP_E_STATE, "not implemented")
int zkp_sha256_transcript(zkp_ctx* c, const uint32_t*, int, int, int, uint8_t*) { NI(c); }
int zkp_rangeproof_ni_prove(zkp_ctx* c, int, int, int, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*,
                            const uint8_t*, const uint32_t*, const uint32_t*, uint32_t*, uint32_t*, uint8_t*, uint8_t*,
                            uint32_t*, uint32_t*) { NI(c); }
int zkp_rp_prove_stage(zkp_ctx* c, int, int, int, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*,
                       const uint8_t*, const uint32_t*, const uint32_t*) { NI(c); }
int zkp_rp_prove_run(zkp_ctx* c) { NI(c); }
int zkp_rp_prove_fetch(zkp_ctx* c, uint32_t*, uint32_t*, uint8_t*, uint8_t*, uint32_t*, uint32_t*) { NI(c); }
int zkp_rangeproof_ni_verify(zkp_ctx* c, int, int, int, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*,
                             const uint8_t*, const uint32_t*, const uint32_t*, uint8_t*, uint8_t*, uint8_t*) { NI(c); }
int zkp_rp_verify_stage(zkp_ctx* c, int, int, int, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*,
                        const uint8_t*, const uint32_t*, const uint32_t*) { NI(c); }
int zkp_rp_verify_stage_from_prove(zkp_ctx* c, const uint32_t*) { NI(c); }
int zkp_rp_verify_run(zkp_ctx* c) { NI(c); }
int zkp_rp_verify_fetch(zkp_ctx* c, uint8_t*, uint8_t*, uint8_t*) { NI(c); }
long long zkp_rp_verify_enc_count(zkp_ctx*) { return 0; }
int zkp_correct_key_ni_verify(zkp_ctx* c, int, int, const uint32_t*, const uint32_t*, const uint8_t*, int, uint8_t*, uint32_t*) { NI(c); }
int zkp_ck_verify_stage(zkp_ctx* c, int, int, const uint32_t*, const uint32_t*, const uint8_t*, int) { NI(c); }
int zkp_ck_verify_run(zkp_ctx* c) { NI(c); }
int zkp_ck_verify_fetch(zkp_ctx* c, uint8_t*, uint32_t*) { NI(c); }
}
