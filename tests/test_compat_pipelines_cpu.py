"""The `diffusers` pipeline stand-ins the reference's example scripts load their base weights from
(flash-diffusion_b200/compat/diffusers): HF-layout checkpoint keys of the four base models, and the SD3 pipeline's
`encode_prompt` (what `FlashDiffusionSD3` conditions on, reference flash_sd3/flash_diffusion_model.py:196-220)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "flash-diffusion_b200", "compat")


@pytest.fixture()
def shim(monkeypatch):
    monkeypatch.setattr(sys, "path", sys.path + [COMPAT])
    for name in [m for m in sys.modules if m == "diffusers" or m.startswith("diffusers.")]:
        monkeypatch.delitem(sys.modules, name)
    import diffusers
    assert diffusers.__version__.endswith("flash_b200_shim")
    yield diffusers
    for name in [m for m in sys.modules if m == "diffusers" or m.startswith("diffusers.")]:
        sys.modules.pop(name, None)


def test_pixart_checkpoint_layout_covers_the_wrapper(shim):
    """examples/train_flash_pixart.py:88-172 loads the HF transformer with strict=False and then copies the
    `adaln_single.emb.*` embedders key by key: every other wrapper parameter must be found under its own name."""
    from flash.models.transformers import DiffusersTransformer2DWrapper
    from flash.recipes import PIXART_KWARGS
    with torch.device("meta"):
        pipe = shim.DiffusionPipeline.from_pretrained("PixArt-alpha/PixArt-XL-2-1024-MS")
        net = DiffusersTransformer2DWrapper(**PIXART_KWARGS)
    hf = pipe.transformer.state_dict()
    own = net.state_dict()
    surgery = ("adaln_single.timestep_embedder.", "adaln_single.add_embedding.")
    for k, v in own.items():
        if not k.startswith(surgery):
            assert k in hf and hf[k].shape == v.shape, k
    for emb, idx in (("resolution_embedder", 0), ("resolution_embedder", 1), ("aspect_ratio_embedder", 2)):
        for lin in ("linear_1", "linear_2"):
            for wb in ("weight", "bias"):
                assert hf[f"adaln_single.emb.{emb}.{lin}.{wb}"].shape == own[f"adaln_single.add_embedding.{idx}.{lin}.{wb}"].shape
    assert hf["adaln_single.emb.timestep_embedder.linear_1.weight"].shape == own["adaln_single.timestep_embedder.linear_1.weight"].shape
    assert not any(k.startswith("adaln_single.add_embedding") for k in hf)
    missing, unexpected = net.load_state_dict(hf, strict=False)
    assert all(k.startswith(surgery) for k in missing) and all(k.startswith("adaln_single.emb.") for k in unexpected)


def test_sd3_and_sdxl_checkpoint_layouts(shim):
    from flash.models.transformers import DiffusersSD3Transformer2DWrapper
    from flash.models.unets import DiffusersUNet2DCondWrapper
    from flash.recipes import SD3_KWARGS, SDXL_UNET_KWARGS
    with torch.device("meta"):
        pipe = shim.StableDiffusion3Pipeline.from_pretrained("stabilityai/stable-diffusion-3-medium",
                                                             text_encoder_3=None, tokenizer_3=None, revision="refs/pr/26")
        assert pipe.text_encoder_3 is None and pipe.text_encoder_2.config.hidden_size == 1280
        net = DiffusersSD3Transformer2DWrapper(**SD3_KWARGS)
        net.load_state_dict(pipe.transformer.state_dict(), strict=True)              # examples/train_flash_sd3.py:79
        xl = shim.StableDiffusionXLPipeline.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0")
        unet = DiffusersUNet2DCondWrapper(**SDXL_UNET_KWARGS)
    hf = xl.unet.state_dict()
    assert "add_embedding.linear_1.weight" in hf and not any(k.startswith("class_embedding") for k in hf)
    missing, unexpected = unet.load_state_dict(hf, strict=False)                      # examples/train_flash_sdxl.py:120-134
    assert all(k.startswith("class_embedding.") for k in missing) and all(k.startswith("add_embedding.") for k in unexpected)
    with pytest.raises(OSError):
        shim.DiffusionPipeline.from_pretrained("someone/some-model")


def _tiny_sd3_pipeline(shim, with_t5):
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection, T5Config, T5EncoderModel

    from flash.models.embedders.offline_text import HashTokenizer
    torch.manual_seed(0)
    pipe = shim.StableDiffusion3Pipeline()
    clip = lambda h, p: CLIPTextModelWithProjection(CLIPTextConfig(
        hidden_size=h, intermediate_size=2 * h, num_hidden_layers=3, num_attention_heads=2, projection_dim=p,
        vocab_size=1000, max_position_embeddings=77)).eval()
    pipe.text_encoder, pipe.text_encoder_2 = clip(16, 12), clip(24, 20)
    pipe.tokenizer = pipe.tokenizer_2 = HashTokenizer(1000, 77, style="clip")
    pipe.text_encoder_3 = pipe.tokenizer_3 = None
    pipe.joint_attention_dim = 64
    if with_t5:
        pipe.text_encoder_3 = T5EncoderModel(T5Config(d_model=64, d_kv=8, d_ff=96, num_layers=2, num_heads=4,
                                                      vocab_size=1000, feed_forward_proj="gated-gelu")).eval()
        pipe.tokenizer_3 = HashTokenizer(1000, 512, style="t5")
    return pipe


@pytest.mark.parametrize("with_t5", [False, True])
def test_sd3_encode_prompt_layout(shim, with_t5):
    """Published `StableDiffusion3Pipeline.encode_prompt`: [CLIP-L | CLIP-G] penultimate states zero-padded to the joint
    width, then the T5 states (77 zero rows without T5: SURVEY §8d config 4, 154 / 333 tokens) on the token axis;
    pooled = the two projections concatenated; negative prompt broadcast over the batch."""
    pipe = _tiny_sd3_pipeline(shim, with_t5)
    prompts = ["a red car on the beach", "a blue pig in space, cartoon"]
    pe, npe, ppe, nppe = pipe.encode_prompt(prompt=prompts, prompt_2=prompts, prompt_3=prompts, negative_prompt="ugly",
                                            negative_prompt_2="ugly", negative_prompt_3="ugly",
                                            do_classifier_free_guidance=True, clip_skip=False, device="cpu")
    T = 77 + (256 if with_t5 else 77)
    assert pe.shape == npe.shape == (2, T, 64) and ppe.shape == nppe.shape == (2, 32)
    # feature layout of the CLIP rows: 16 + 24 real columns, zero padding up to the joint width
    assert pe[:, :77, :40].abs().sum() > 0 and pe[:, :77, 40:].abs().sum() == 0
    assert (pe[:, 77:].abs().sum() > 0) == with_t5
    assert torch.equal(npe[0], npe[1]) and not torch.equal(pe[0], pe[1])
    # penultimate hidden state (clip_skip=False counts as 0 -> hidden_states[-2]) and projected pooled embedding
    tok = pipe.tokenizer(prompts, padding="max_length", max_length=77, truncation=True, return_tensors="pt")
    ref = pipe.text_encoder(tok["input_ids"], output_hidden_states=True)
    assert torch.allclose(pe[:, :77, :16], ref.hidden_states[-2], atol=1e-6)
    assert torch.allclose(ppe[:, :12], ref.text_embeds, atol=1e-6)


def test_sd3_model_conditions_on_the_pipeline(shim):
    """FlashDiffusionSD3 with `pipeline=` (how examples/train_flash_sd3.py:204-216 builds it): the four tensors of
    `encode_prompt` become the cond / uncond `vector` + `crossattn`."""
    from flash.models.flash_sd3 import FlashDiffusionSD3, FlashDiffusionSD3Config
    from flash.schedulers import FlowMatchEulerDiscreteScheduler
    pipe = _tiny_sd3_pipeline(shim, False)
    sched = FlowMatchEulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-3-medium", subfolder="scheduler")
    model = FlashDiffusionSD3(FlashDiffusionSD3Config(K=[4], num_iterations_per_K=[10]), student_denoiser=torch.nn.Identity(),
                              teacher_denoiser=torch.nn.Identity(), teacher_noise_scheduler=sched,
                              sampling_noise_scheduler=sched, pipeline=pipe)
    cond, uncond = model._conditionings({"text": ["a", "b c"]}, torch.device("cpu"))
    assert cond["cond"]["crossattn"].shape == uncond["cond"]["crossattn"].shape == (2, 154, 64)
    assert cond["cond"]["vector"].shape == uncond["cond"]["vector"].shape == (2, 32)
