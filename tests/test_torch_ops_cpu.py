"""The reference's native operator schema (flash_api.cpp:275-365) is registered under torch.ops.hstu with identical signatures."""
import torch


def test_schemas_match_the_reference_declarations():
    from generative_recommenders_b200 import torch_ops

    torch_ops.register()
    torch_ops.register()  # idempotent
    s = str(torch.ops.hstu.hstu_mha.default._schema)
    assert s.startswith("hstu::hstu_mha(SymInt max_seq_len, float alpha, Tensor q, Tensor k, Tensor v, Tensor? seq_offsets, bool causal, "
                        "Tensor? num_targets, Tensor? attn_scale, int max_attn_len, int min_full_attn_seq_len, int contextual_seq_len, "
                        "Tensor? q_descale, Tensor? k_descale, Tensor? v_descale, bool sort_by_length, bool deterministic, int sm_margin)")
    assert "hstu_mha_fwd(SymInt max_seq_len" in str(torch.ops.hstu.hstu_mha_fwd.default._schema)
    b = str(torch.ops.hstu.hstu_mha_bwd.default._schema)
    assert "Tensor dq, Tensor dk, Tensor dv" in b and b.endswith("-> Tensor[]")
    # shape inference without a device
    q = torch.empty(10, 2, 32, device="meta", dtype=torch.bfloat16)
    v = torch.empty(10, 2, 64, device="meta", dtype=torch.bfloat16)
    off = torch.empty(3, device="meta", dtype=torch.int32)
    out = torch.ops.hstu.hstu_mha_fwd(16, 0.1, q, q, v, off, True, None, None, 0, 0, 0, None, None, None, 0)
    assert out.shape == (10, 2, 64) and out.dtype == torch.bfloat16
