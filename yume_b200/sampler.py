"""Euler / SDE denoising loops of the Yume samplers — the callers of `WanModel.forward`.

The reference runs these loops inline in its scripts:
  5B   fastvideo/sample/sample_5b.py:941-1034   `sample_step` Euler steps, shift 7.0, one forward per step, history frames
       are clean latents (per-token t: 0 for history tokens, sigma*1000 for the new ones)
  14B  fastvideo/sample/sample.py:755-790       Euler + classifier-free guidance 5.0 (cond + uncond forward per step),
       history frames re-noised to the next sigma every step
  SDE  fastvideo/sample/sample_tts.py:726-744   stochastic update (eta 0.3) applied on top of the Euler step
They are restated here as functions over ANY callable with the reference's `WanModel.forward` signature (the reference
module with `yume_b200.install()` applied, or the mirrors in yume_b200.model), so bench.py can time BASELINE.json
configs[2] / configs[3] as loops. What makes the loop fast lives in the engine, behind the unchanged forward signature:
the text/CLIP embedding and every block's cross-attention K/V depend only on the context, so the engine keeps them
across steps (WanDiT context cache), and the fixed-shape step can be replayed as one CUDA graph (WanDiT.use_cuda_graph).
The arithmetic of the update itself is a few elementwise torch ops on a 14 MB latent — host-side glue, as in the reference.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch

Tensor = torch.Tensor


def sampling_sigmas(steps: int, shift: float) -> list:
    """get_sampling_sigmas (sample_5b.py:502-506): sigma_i = shift*s/(1+(shift-1)*s), s = linspace(1, 0, steps+1)[:steps]."""
    out = []
    for i in range(steps):
        s = 1.0 - i / steps
        out.append(shift * s / (1.0 + (shift - 1.0) * s))
    return out


def sde_update(latent_new: Tensor, noise_pred: Tensor, temp_x0: Tensor, sigma: float, sigma_next: float, last: bool,
               eta: float = 0.3, generator: Optional[torch.Generator] = None) -> Tensor:
    """sample_tts.py:726-744 — stochastic correction of the Euler result `temp_x0` for the new frames.
    `last` is the reference's `i + 1 == 50` test (its schedule has 50 steps)."""
    pred_original = latent_new + (0.0 - sigma) * noise_pred
    delta_t = 0.0 if last else max(sigma - sigma_next, 0.0)
    dsigma = (0.0 - sigma) if last else (sigma_next - sigma)
    std_dev_t = eta * math.sqrt(delta_t)
    score = -(latent_new - pred_original * (1.0 - sigma)) / (sigma ** 2)
    mean = temp_x0 + (-0.5 * eta ** 2 * score) * dsigma
    noise = torch.randn(mean.shape, device=mean.device, dtype=mean.dtype, generator=generator)
    return mean + noise * std_dev_t


@torch.no_grad()
def denoise_chunk_5b(transformer: Callable, latent: Tensor, history: Tensor, latent_frame_zero: int, steps: int,
                     arg_c: Dict, shift: float = 7.0, sde: bool = False, eta: float = 0.3,
                     generator: Optional[torch.Generator] = None) -> Tensor:
    """One autoregressive chunk of the 5B sampler (sample_5b.py:941-1034, the `step_sample > 0` / i2v branch).
    latent  [48, H_f + lfz, h, w]: clean history frames followed by noise for the `latent_frame_zero` new frames
    history [48, H_f, h, w]: the clean history latents (`model_input_1[:, :-lfz]`), re-attached after every step
    arg_c   the reference's kwargs (`context` list, `seq_len`)
    Returns the denoised latent [48, H_f + lfz, h, w]. `sde` applies sample_tts.py's update to the 5B loop (BASELINE.json
    configs[3]: "Yume-5B SDE/TTS 4-step distilled sampling")."""
    lfz = latent_frame_zero
    sig = sampling_sigmas(steps, shift)
    for i in range(steps):
        t = torch.tensor([[0.0, sig[i] * 1000.0]], device=latent.device)          # history tokens t = 0, new tokens sigma*1000
        pred = transformer([latent], t=t, latent_frame_zero=lfz, **arg_c)[0]
        pred = pred[:, -lfz:]
        nxt = 0.0 if i + 1 == steps else sig[i + 1]
        new = latent[:, -lfz:]
        x0 = new + (nxt - sig[i]) * pred
        if sde:
            x0 = sde_update(new, pred, x0, sig[i], nxt, i + 1 == steps, eta, generator)
        latent = torch.cat([history, x0], dim=1)
    return latent


@torch.no_grad()
def denoise_chunk_14b(transformer: Callable, noise: Tensor, model_input: Tensor, latent_frame_zero: int, steps: int,
                      arg_c: Dict, arg_null: Dict, rand_num_img: float = 0.6, shift: float = 3.0, guidance: float = 5.0,
                      sde: bool = False, eta: float = 0.3, generator: Optional[torch.Generator] = None,
                      first_steps: Optional[int] = None) -> Tensor:
    """One chunk of the 14B sampler (sample.py:755-790): Euler ODE with classifier-free guidance — two forwards per
    step with the SAME latent and timestep, contexts `arg_c` / `arg_null`; the history frames are re-noised to the next
    sigma after every step. `first_steps` runs only the first k steps of the `steps`-step schedule (bench slices)."""
    lfz = latent_frame_zero
    sig = sampling_sigmas(steps, shift)
    latent = noise
    for i in range(steps if first_steps is None else min(first_steps, steps)):
        t = torch.tensor([sig[i] * 1000.0], device=latent.device)
        cond, _ = transformer([latent], t=t, rand_num_img=rand_num_img, latent_frame_zero=lfz, **arg_c)
        uncond, _ = transformer([latent], t=t, rand_num_img=rand_num_img, latent_frame_zero=lfz, **arg_null)
        pred = uncond + guidance * (cond - uncond)
        pred = pred[:, -lfz:]
        nxt = 0.0 if i + 1 == steps else sig[i + 1]
        new = latent[:, -lfz:]
        x0 = new + (nxt - sig[i]) * pred
        if sde:
            x0 = sde_update(new, pred, x0, sig[i], nxt, i + 1 == steps, eta, generator)
        s1 = sig[min(steps - 1, i + 1)]
        latent = torch.cat([noise[:, :-lfz] * s1 + (1.0 - s1) * model_input[:, :-lfz], x0], dim=1)
    return latent
