"""Host-logic tests (no GPU) of the dual-stream family mirrors, the HunyuanVideo-1.5 VAE mirror and the CUDA-graph helper:
configuration validation, tile grids, state-dict layout conversion, error behaviour (the product path has no CPU fallback)."""
import pytest
import torch

import hy15_vae


def test_graphed_callable_needs_cuda():
    from apex_studio_b200.graph import GraphedCallable

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        GraphedCallable(lambda x: x, (torch.zeros(2),))


def test_hy15_vae_config_tile_grid_and_weight_layouts():
    from apex_studio_b200.vae import AutoencoderKLHunyuanVideo15, HunyuanVideo15VAEConfig

    vae = AutoencoderKLHunyuanVideo15()
    assert vae.dims == [1024, 1024, 512, 256, 128] and vae.tile_latent_min_height == 8
    assert [vae._up_flags(i) for i in range(5)] == [(True, True), (True, True), (True, False), (True, False), (False, False)]
    grid = vae.tile_grid(45, 80)                                  # 720p: 8 x 14 = 112 tiles of 8x8 latents at stride 6
    assert len(grid) == 112 and grid[0] == (0, 0) and grid[-1] == (42, 78) and grid[14] == (6, 0)
    with pytest.raises(ValueError):
        AutoencoderKLHunyuanVideo15(HunyuanVideo15VAEConfig(block_out_channels=(48, 96, 192, 384, 384)))   # not multiples of 32
    with pytest.raises(ValueError):
        vae.enable_tiling(use_light_vae=True)
    with pytest.raises(RuntimeError):
        vae.decode(torch.zeros(1, 32, 1, 8, 8))                 # weights not loaded
    # state-dict conversion on the CPU: tap-major conv weights, fused q|k, flat gammas, conv_out padded to 16 channels
    ch = (64, 64, 32, 32, 32)
    small = AutoencoderKLHunyuanVideo15(HunyuanVideo15VAEConfig(block_out_channels=tuple(reversed(ch))))
    w = hy15_vae.make_weights(ch, seed=1)
    small.load_state_dict(dict(w, **{"encoder.conv_in.conv.weight": torch.zeros(1)}), device="cpu")
    assert small.w["decoder.conv_in.conv.weight"].shape == (27 * 64, 32)
    assert small.w["decoder.conv_out.conv.weight"].shape == (27 * 16, 32) and small.w["decoder.conv_out.conv.bias"].shape == (16,)
    assert small.w["decoder.mid_block.attentions.0.to_qk.weight"].shape == (128, 64)
    assert small.w["decoder.mid_block.resnets.0.norm1.gamma"].shape == (64,)
    assert not any(k.startswith("encoder.") for k in small.w)
    tap0 = w["decoder.conv_in.conv.weight"][:, :, 0, 0, 0].bfloat16()           # tap (0,0,0) block = rows [0, Cout)
    assert torch.equal(small.w["decoder.conv_in.conv.weight"][:64], tap0)
    with pytest.raises(ValueError):                                               # no CPU fallback
        small.decode(torch.zeros(1, 32, 1, 8, 8))
    lat = torch.randn(1, 32, 2, 4, 4)
    assert torch.allclose(small.denormalize_latents(lat), lat / 1.03682)


def test_qwen_forward_argument_validation_and_config_flags():
    from apex_studio_b200.qwenimage import QwenImageConfig, QwenImageTransformer2DModel

    for flag in ("zero_cond_t", "use_additional_t_cond", "use_layer3d_rope", "guidance_embeds"):
        with pytest.raises(ValueError):
            QwenImageTransformer2DModel(QwenImageConfig(**{flag: True}))
    with pytest.raises(ValueError):
        QwenImageTransformer2DModel(QwenImageConfig(attention_head_dim=64))
    m = QwenImageTransformer2DModel(QwenImageConfig(num_layers=1, num_attention_heads=2, joint_attention_dim=48, in_channels=16, out_channels=4))
    with pytest.raises(RuntimeError):
        m(hidden_states=torch.zeros(1, 4, 16), encoder_hidden_states=torch.zeros(1, 2, 48), timestep=torch.zeros(1), img_shapes=[[(1, 2, 2)]],
          txt_seq_lens=[2])


def test_flux_family_config_validation():
    from apex_studio_b200.flux import FluxConfig, FluxTransformer2DModel
    from apex_studio_b200.flux2 import Flux2Config, Flux2Transformer2DModel
    from apex_studio_b200.hunyuanvideo15 import HunyuanVideo15Config, HunyuanVideo15Transformer3DModel

    with pytest.raises(ValueError):
        FluxTransformer2DModel(FluxConfig(axes_dims_rope=(16, 48, 48)))
    with pytest.raises(ValueError):
        FluxTransformer2DModel(FluxConfig(patch_size=2))
    with pytest.raises(ValueError):
        Flux2Transformer2DModel(Flux2Config(axes_dims_rope=(32, 32, 32)))
    with pytest.raises(ValueError):
        HunyuanVideo15Transformer3DModel(HunyuanVideo15Config(use_meanflow=True))
    m = HunyuanVideo15Transformer3DModel.from_config(dict(num_layers=2, num_attention_heads=2, rope_axes_dim=[16, 56, 56], unknown_key=1))
    assert m.config.num_layers == 2 and m.config.rope_axes_dim == (16, 56, 56)
    assert Flux2Transformer2DModel(Flux2Config(num_attention_heads=24)).mlp == 9216
    f = FluxTransformer2DModel(FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2))
    assert [n for n, _ in f._modulation_layout()] == ["transformer_blocks.0.norm1.linear", "transformer_blocks.0.norm1_context.linear",
                                                      "single_transformer_blocks.0.norm.linear", "norm_out.linear"]
