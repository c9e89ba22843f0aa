// placeholder until the dense path lands (next commit): every entry point reports UNSUPPORTED.
#include "common.cuh"
#define STUB(name, ...) extern "C" __attribute__((visibility("default"))) int name(__VA_ARGS__) { bdetr::set_error(#name ": not implemented yet"); return BDETR_E_UNSUPPORTED; }
STUB(bdetr_attention_block_fwd, int, int, int, int, int, const float *, const float *, const float *, const bdetr_attn_params *, float, uint32_t, float, float *, const bdetr_attn_saved *, void *)
STUB(bdetr_attention_block_bwd, int, int, int, int, int, const float *, const float *, const float *, const bdetr_attn_params *, float, uint32_t, const bdetr_attn_saved *, const float *, float *, float *, float *, int, const bdetr_attn_params *, const bdetr_attn_scratch *, void *)
STUB(bdetr_ffn_block_fwd, int, int, const float *, const bdetr_ffn_params *, float, uint32_t, float, float *, const bdetr_ffn_saved *, void *)
STUB(bdetr_ffn_block_bwd, int, int, const float *, const bdetr_ffn_params *, float, uint32_t, const bdetr_ffn_saved *, const float *, float *, int, const bdetr_ffn_params *, const bdetr_ffn_scratch *, void *)
STUB(bdetr_add_positional_fwd, int, int, int, const float *, const float *, float *, void *)
STUB(bdetr_add_positional_bwd, int, int, int, const float *, float *, void *)
STUB(bdetr_tile_queries_fwd, int, int, int, const float *, float *, void *)
STUB(bdetr_accumulate, size_t, const float *, float *, void *)
STUB(bdetr_head_fwd, int, int, int, int, int, int, float, const float *, const bdetr_head_params *, float, float, float *, int, const bdetr_head_saved *, void *)
STUB(bdetr_head_bwd, int, int, int, int, int, float, const float *, const bdetr_head_params *, float, const bdetr_head_saved *, const float *, float *, int, const bdetr_head_params *, const bdetr_head_scratch *, void *)
STUB(bdetr_gemm, int, int, int, const float *, int, const float *, int, const float *, int, int, float *, void *)
