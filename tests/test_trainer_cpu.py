"""The drop-in's way of getting final hidden states out of an UNMODIFIED Hugging Face model
(`trainer.final_hidden_states`: lm_head tapped for one forward) on the real Qwen2.5-VL class with video
kwargs, exactly as the reference calls `model(input_ids, **prompt_inputs)` (grpo_trainer.py:375, :603-611).
Pure torch, runs on the CPU box (tiny random-init config)."""
import pytest
import torch

tf = pytest.importorskip("transformers")


def test_lm_head_tap_on_real_qwen2_5_vl_with_video_inputs():
    if not hasattr(tf, "Qwen2_5_VLForConditionalGeneration"):
        pytest.skip("transformers without Qwen2.5-VL")
    from open_o3_video_b200.trainer import final_hidden_states, lm_head_weight
    cfg = tf.Qwen2_5_VLConfig(
        text_config=dict(vocab_size=600, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, max_position_embeddings=256,
                         rope_scaling={"type": "mrope", "mrope_section": [2, 3, 3]}),
        vision_config=dict(depth=1, hidden_size=32, intermediate_size=64, num_heads=2, out_hidden_size=64,
                           fullatt_block_indexes=[0], window_size=56),
        video_token_id=599, image_token_id=598, vision_start_token_id=597, vision_end_token_id=596)
    torch.manual_seed(0)
    model = tf.Qwen2_5_VLForConditionalGeneration(cfg).eval()
    vc = cfg.vision_config
    ids = torch.randint(0, 500, (2, 12))
    ids[:, 2], ids[:, 3:7], ids[:, 7] = 597, 599, 596                   # <vision_start> 4 video tokens <vision_end>
    kwargs = dict(pixel_values_videos=torch.randn(2 * 16, 3 * vc.temporal_patch_size * vc.patch_size ** 2),
                  video_grid_thw=torch.tensor([[1, 4, 4], [1, 4, 4]]), attention_mask=torch.ones_like(ids))
    with torch.no_grad():
        logits = model(ids, **kwargs).logits                            # what the reference materialises
        hidden = final_hidden_states(model, ids, **kwargs)
        text_only = final_hidden_states(model, ids)
    assert isinstance(model.lm_head, torch.nn.Linear)                   # the head is back in place
    assert hidden.shape == (2, 12, 64)
    assert torch.allclose(hidden.float() @ lm_head_weight(model).float().T, logits.float(), atol=1e-4)
    assert not torch.allclose(hidden, text_only)                        # the vision tower really ran


def test_lm_head_tap_restores_the_head_on_error():
    from open_o3_video_b200.trainer import final_hidden_states

    class Broken(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lm_head = torch.nn.Linear(4, 8, bias=False)

        def forward(self, input_ids, **kw):
            raise ValueError("boom")

    m = Broken()
    with pytest.raises(ValueError):
        final_hidden_states(m, torch.zeros(1, 2, dtype=torch.long))
    assert isinstance(m.lm_head, torch.nn.Linear)
