// placeholder, replaced below
#include "../../include/b200w.h"
#include "common.h"
using namespace b200w;
#define NI(name) { set_last_error(name " not implemented yet"); return -1; }
extern "C" {
int b200w_engine_create(const b200w_model_config*, b200w_engine**) NI("engine_create")
int b200w_engine_destroy(b200w_engine*) NI("engine_destroy")
int b200w_engine_set_weight(b200w_engine*, const char*, const void*, int64_t, int) NI("set_weight")
int b200w_engine_finalize(b200w_engine*) NI("finalize")
int b200w_encode(b200w_engine*, const void*, int, void*, void*) NI("encode")
int b200w_cross_kv(b200w_engine*, const void*, int, void*, void*) NI("cross_kv")
int b200w_decode(b200w_engine*, const void*, int, const int32_t*, const int32_t*, const b200w_decode_opts*, int32_t*, int32_t, int32_t*, float*, float*, void*) NI("decode")
int b200w_align(b200w_engine*, const void*, int, const int32_t*, const int32_t*, const int32_t*, int32_t, const int32_t*, int32_t, int32_t, int32_t*, int32_t, int32_t*, float*, int32_t, void*) NI("align")
int b200w_detect_language(b200w_engine*, const void*, int, int32_t, const int32_t*, int32_t, float*, void*) NI("detect_language")
int b200w_decoder_logits(b200w_engine*, const void*, int, const int32_t*, const int32_t*, float*, void*) NI("decoder_logits")
}
