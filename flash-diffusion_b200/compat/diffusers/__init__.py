"""Shim of the `diffusers` names the Flash-Diffusion example scripts import (see ../README.md)."""
import torch

from flash.schedulers import (DDPMScheduler, DPMSolverMultistepScheduler, EulerAncestralDiscreteScheduler,  # noqa: F401
                              EulerDiscreteScheduler, FlashFlowMatchEulerDiscreteScheduler,
                              FlowMatchEulerDiscreteScheduler, LCMScheduler)

__version__ = "0.0.0+flash_b200_shim"


class _Weights:
    """stands in for `pipe.unet`: a random-init state dict in the HF key layout of the named checkpoint"""

    def __init__(self, sd):
        self._sd = sd

    def state_dict(self):
        return self._sd


def _materialise(net, seed):
    """random-init weights on the CPU — unless the caller runs under `torch.device("meta")` (the plumbing tests of the
    full-size scripts: shapes and key names only, no memory touched)"""
    from flash.recipes import init_random_
    if torch.get_default_device().type == "meta":
        return net
    return init_random_(net.to_empty(device="cpu"), seed)


def _sd15_unet_state_dict(seed=0):
    """`runwayml/stable-diffusion-v1-5` UNet keys: as DiffusersUNet2DCondWrapper's (examples/train_flash_sd.py:56-114)
    except that the Transformer2D proj_in / proj_out are 1x1 convolutions ([C, C, 1, 1]) — the script squeezes them
    (:116-158)."""
    from flash.models.unets import DiffusersUNet2DCondWrapper
    from flash.recipes import SD15_UNET_KWARGS, init_random_
    with torch.device("meta"):
        net = DiffusersUNet2DCondWrapper(**dict(SD15_UNET_KWARGS, use_linear_projection=True))
    net = _materialise(net, 1234 + seed)
    sd = {}
    for k, v in net.state_dict().items():
        if k.endswith(("proj_in.weight", "proj_out.weight")) and ".attentions." in k:
            v = v[:, :, None, None]
        sd[k] = v
    return sd


def _sdxl_unet_state_dict(seed=0):
    """`stabilityai/stable-diffusion-xl-base-1.0` UNet keys: the wrapper's, with the vector conditioning under
    `add_embedding.*` (the script copies it into `class_embedding`, examples/train_flash_sdxl.py:120-134)."""
    from flash.models.unets import DiffusersUNet2DCondWrapper
    from flash.recipes import SDXL_UNET_KWARGS, init_random_
    with torch.device("meta"):
        net = DiffusersUNet2DCondWrapper(**SDXL_UNET_KWARGS)
    net = _materialise(net, 1234 + seed)
    sd = {}
    for k, v in net.state_dict().items():
        sd[k.replace("class_embedding.", "add_embedding.")] = v
    return sd


def _pixart_transformer_state_dict(seed=0):
    """`PixArt-alpha/PixArt-XL-2-1024-MS` transformer keys: the wrapper's, except that the HF checkpoint keeps the
    micro-conditioning under `adaln_single.emb.{timestep,resolution,aspect_ratio}_embedder` with ONE resolution embedder
    shared by height and width — the script copies it into `add_embedding[0]` AND `[1]`, the aspect-ratio embedder into
    `[2]` (examples/train_flash_pixart.py:90-172)."""
    from flash.models.transformers import DiffusersTransformer2DWrapper
    from flash.recipes import PIXART_KWARGS
    with torch.device("meta"):
        net = DiffusersTransformer2DWrapper(**PIXART_KWARGS)
    net = _materialise(net, 1234 + seed)
    sd = {}
    for k, v in net.state_dict().items():
        if k.startswith("adaln_single.add_embedding.1."):
            continue
        k = k.replace("adaln_single.timestep_embedder.", "adaln_single.emb.timestep_embedder.")
        k = k.replace("adaln_single.add_embedding.0.", "adaln_single.emb.resolution_embedder.")
        k = k.replace("adaln_single.add_embedding.2.", "adaln_single.emb.aspect_ratio_embedder.")
        sd[k] = v
    return sd


def _sd3_transformer_state_dict(seed=0):
    """`stabilityai/stable-diffusion-3-medium` transformer keys = `SD3Transformer2DModel`'s = the wrapper's (the script
    loads them with strict=True, examples/train_flash_sd3.py:79)."""
    from flash.models.transformers import DiffusersSD3Transformer2DWrapper
    from flash.recipes import SD3_KWARGS
    with torch.device("meta"):
        net = DiffusersSD3Transformer2DWrapper(**SD3_KWARGS)
    net = _materialise(net, 1234 + seed)
    return dict(net.state_dict())


class DiffusionPipeline:
    _BUILDERS = {"runwayml/stable-diffusion-v1-5": ("unet", _sd15_unet_state_dict),
                 "stabilityai/stable-diffusion-xl-base-1.0": ("unet", _sdxl_unet_state_dict),
                 "PixArt-alpha/PixArt-XL-2-1024-MS": ("transformer", _pixart_transformer_state_dict),
                 "stabilityai/stable-diffusion-3-medium": ("transformer", _sd3_transformer_state_dict)}

    @classmethod
    def from_pretrained(cls, repo, **kwargs):
        if repo not in cls._BUILDERS:
            raise OSError(f"{repo}: no network and no local copy; the offline shim builds random-init weights for "
                          f"{sorted(cls._BUILDERS)} only")
        attr, build = cls._BUILDERS[repo]
        pipe = cls()
        setattr(pipe, attr, _Weights(build()))
        pipe._after_load(repo, **kwargs)
        return pipe

    def _after_load(self, repo, **kwargs):
        pass

    def to(self, *args, **kwargs):
        return self


class StableDiffusionXLPipeline(DiffusionPipeline):
    pass


class StableDiffusion3Pipeline(DiffusionPipeline):
    """What `FlashDiffusionSD3` uses of the diffusers SD3 pipeline (src/flash/models/flash_sd3/flash_diffusion_model.py
    :196-220, :715-736): `.transformer` / `.vae` (deleted by the script after the weights are taken), `.to(device)` and
    `.encode_prompt(...)`.  `encode_prompt` restates the published pipeline (diffusers 0.29
    `StableDiffusion3Pipeline.encode_prompt`): CLIP-L and OpenCLIP-bigG penultimate hidden states concatenated on the
    feature axis and zero-padded to the T5 width, then the T5-XXL states (or 77 zero rows when the pipeline was loaded
    with `text_encoder_3=None`) appended on the token axis; pooled = the two projected CLIP embeddings concatenated.
    The encoders are `transformers`' own modules (random weights offline), as in the reference."""

    tokenizer_max_length = 77
    joint_attention_dim = 4096
    _REPO = "stabilityai/stable-diffusion-3-medium"

    def _after_load(self, repo, text_encoder_3="default", tokenizer_3="default", revision=None, **unused):
        from transformers import CLIPTextModelWithProjection, CLIPTokenizer, T5EncoderModel, T5Tokenizer

        from flash.models.embedders.offline_text import load_text_model, load_tokenizer
        self.vae = None           # the script deletes it and builds its own AutoencoderKLDiffusers (:82-96)
        self.text_encoder = load_text_model(CLIPTextModelWithProjection, repo, "text_encoder", revision).eval()
        self.text_encoder_2 = load_text_model(CLIPTextModelWithProjection, repo, "text_encoder_2", revision).eval()
        self.tokenizer = load_tokenizer(CLIPTokenizer, repo, "tokenizer", revision, "clip")
        self.tokenizer_2 = load_tokenizer(CLIPTokenizer, repo, "tokenizer_2", revision, "clip")
        self.text_encoder_3 = self.tokenizer_3 = None
        if text_encoder_3 is not None:
            self.text_encoder_3 = load_text_model(T5EncoderModel, repo, "text_encoder_3", revision).eval()
            self.tokenizer_3 = load_tokenizer(T5Tokenizer, repo, "tokenizer_3", revision, "t5")
        for enc in (self.text_encoder, self.text_encoder_2, self.text_encoder_3):
            if enc is not None:
                enc.requires_grad_(False)

    def to(self, *args, **kwargs):
        for name in ("text_encoder", "text_encoder_2", "text_encoder_3"):
            enc = getattr(self, name, None)
            if enc is not None:
                enc.to(*args, **kwargs)
        return self

    def _clip(self, prompts, tokenizer, encoder, device, clip_skip):
        tok = tokenizer(prompts, padding="max_length", max_length=self.tokenizer_max_length, truncation=True,
                        return_tensors="pt")
        out = encoder(tok["input_ids"].to(device), output_hidden_states=True)
        layer = -2 if clip_skip is None else -(int(clip_skip) + 2)
        return out.hidden_states[layer], out.text_embeds

    def _t5(self, prompts, device, max_sequence_length, dtype):
        if self.text_encoder_3 is None:
            return torch.zeros(len(prompts), self.tokenizer_max_length, self.joint_attention_dim, device=device,
                               dtype=dtype)
        tok = self.tokenizer_3(prompts, padding="max_length", max_length=max_sequence_length, truncation=True,
                               add_special_tokens=True, return_tensors="pt")
        return self.text_encoder_3(tok["input_ids"].to(device))[0].to(dtype)

    def _embed(self, p1, p2, p3, device, clip_skip, max_sequence_length):
        h1, pooled1 = self._clip(p1, self.tokenizer, self.text_encoder, device, clip_skip)
        h2, pooled2 = self._clip(p2, self.tokenizer_2, self.text_encoder_2, device, clip_skip)
        clip = torch.cat([h1, h2], dim=-1)
        t5 = self._t5(p3, device, max_sequence_length, clip.dtype)
        clip = torch.nn.functional.pad(clip, (0, t5.shape[-1] - clip.shape[-1]))
        return torch.cat([clip, t5], dim=-2), torch.cat([pooled1, pooled2], dim=-1)

    @torch.no_grad()
    def encode_prompt(self, prompt, prompt_2=None, prompt_3=None, device=None, num_images_per_prompt=1,
                      do_classifier_free_guidance=True, negative_prompt=None, negative_prompt_2=None,
                      negative_prompt_3=None, prompt_embeds=None, negative_prompt_embeds=None,
                      pooled_prompt_embeds=None, negative_pooled_prompt_embeds=None, clip_skip=None,
                      max_sequence_length=256, **unused):
        as_list = lambda p: [p] if isinstance(p, str) else list(p)
        prompt = as_list(prompt)
        n = len(prompt)
        device = device if device is not None else next(self.text_encoder.parameters()).device
        if prompt_embeds is None:
            prompt_embeds, pooled_prompt_embeds = self._embed(
                prompt, as_list(prompt_2 or prompt), as_list(prompt_3 or prompt), device, clip_skip, max_sequence_length)
        if do_classifier_free_guidance and negative_prompt_embeds is None:
            bcast = lambda p: (n * [p] if isinstance(p, str) else list(p))
            neg = bcast(negative_prompt or "")
            if len(neg) != n:
                raise ValueError(f"`negative_prompt` has batch size {len(neg)}, `prompt` has {n}")
            negative_prompt_embeds, negative_pooled_prompt_embeds = self._embed(
                neg, bcast(negative_prompt_2 or negative_prompt or ""), bcast(negative_prompt_3 or negative_prompt or ""),
                device, None, max_sequence_length)
        if num_images_per_prompt != 1:
            rep = lambda t: None if t is None else t.repeat_interleave(num_images_per_prompt, dim=0)
            prompt_embeds, negative_prompt_embeds = rep(prompt_embeds), rep(negative_prompt_embeds)
            pooled_prompt_embeds, negative_pooled_prompt_embeds = rep(pooled_prompt_embeds), rep(negative_pooled_prompt_embeds)
        return prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds
