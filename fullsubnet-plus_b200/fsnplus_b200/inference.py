"""Batched / multi-GPU inference harness around the model forward.

Restates the 10 lines of the reference inferencer methods that surround the model call
(speech_enhance/fullsubnet_plus/inferencer/inferencer.py:140-165 ``mag_complex_full_band_crm_mask`` and
:116-137 ``full_band_crm_mask``): STFT -> model -> decompress_cIRM -> complex multiply -> iSTFT, with STFT and
iSTFT kept in PyTorch on the GPU as BASELINE.json's north_star prescribes, for a BATCH of equal-length clips
(the reference is hard-wired to batch 1, audio_zen/inferencer/base_inferencer.py:65-69).

Multi-GPU: utterances are independent, so a batch is sharded across ranks (one process per GPU) with no
data-path collective; the only exchange is one all-gather of the enhanced waveforms (NCCL on GPUs, gloo in
the CPU tests of the host logic).
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def stft(y, n_fft=512, hop_length=256, win_length=512):
    """reference audio_zen/acoustics/feature.py:10-31.  y [B, L] -> complex [B, F, T]."""
    assert y.dim() == 2
    return torch.stft(y, n_fft, hop_length, win_length, window=torch.hann_window(n_fft, device=y.device),
                      return_complex=True)


def istft(spec, n_fft=512, hop_length=256, win_length=512, length=None):
    """reference audio_zen/acoustics/feature.py:34-65 (complex input)."""
    return torch.istft(spec, n_fft, hop_length, win_length, window=torch.hann_window(n_fft, device=spec.device),
                       length=length)


def decompress_cIRM(mask, K=10, limit=9.9):
    """reference audio_zen/acoustics/mask.py:60-63."""
    mask = limit * (mask >= limit) - limit * (mask <= -limit) + mask * (torch.abs(mask) < limit)
    return -K * torch.log((K - mask) / (K + mask))


def apply_cirm(crm, X):
    """decompress_cIRM + complex multiply (reference inferencer.py:152-157): one fused CUDA kernel behind the C ABI for
    CUDA float32 inputs, the torch restatement otherwise (CPU tests of the harness)."""
    if crm.is_cuda and crm.dtype == torch.float32 and X.dtype == torch.complex64:
        crm = crm.contiguous()
        Xr = torch.view_as_real(X.contiguous())
        out = torch.empty_like(Xr)
        B, _, F, T = crm.shape
        with torch.cuda.device(crm.device):
            stream = C.c_void_p(torch.cuda.current_stream(crm.device).cuda_stream)
            _lib.check(_lib.load_library().fsn_apply_cirm(C.c_void_p(crm.data_ptr()), C.c_void_p(Xr.data_ptr()), C.c_void_p(out.data_ptr()),
                                                          B, F, T, stream))
        return torch.view_as_complex(out)
    m = decompress_cIRM(crm)                                   # [B, 2, F, T]
    return torch.complex(m[:, 0] * X.real - m[:, 1] * X.imag, m[:, 1] * X.real + m[:, 0] * X.imag)


@torch.no_grad()
def enhance_batch(model, noisy, n_fft=512, hop_length=256, win_length=512, complex_inputs=True):
    """noisy [B, L] float32 on the model's device -> enhanced [B, L].
    complex_inputs=True mirrors mag_complex_full_band_crm_mask (FullSubNet_Plus), False full_band_crm_mask (Model)."""
    X = stft(noisy, n_fft, hop_length, win_length)
    mag = X.abs().unsqueeze(1)
    if complex_inputs:
        crm = model(mag, X.real.unsqueeze(1).contiguous(), X.imag.unsqueeze(1).contiguous())
    else:
        crm = model(mag)
    return istft(apply_cirm(crm, X), n_fft, hop_length, win_length, length=noisy.size(-1))


def shard_range(n_items, rank, world_size):
    """Contiguous shard [lo, hi) of rank; sizes differ by at most one (ragged batches are allowed)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_enhanced(local, n_items, group=None):
    """Gather per-rank enhanced waveforms [n_local, L] into [n_items, L] on every rank with ONE collective
    (all_gather_into_tensor on equal shards, padded all_gather when the batch is ragged)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    L = local.size(1)
    per = (n_items + world - 1) // world
    buf = local
    if local.size(0) != per:
        buf = torch.zeros((per, L), dtype=local.dtype, device=local.device)
        buf[: local.size(0)] = local
    out = torch.empty((world * per, L), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    if per * world == n_items:
        return out
    pieces = []
    for r in range(world):
        lo, hi = shard_range(n_items, r, world)
        pieces.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(pieces, 0)


@torch.no_grad()
def enhance_sharded(model, noisy_all, group=None, **kw):
    """Data-parallel enhancement: every rank holds the full list of clips (or at least its shard), enhances its
    contiguous shard and all-gathers the waveforms.  noisy_all [N, L] on the local device."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_range(noisy_all.size(0), rank, world)
    local = enhance_batch(model, noisy_all[lo:hi], **kw) if hi > lo else noisy_all.new_zeros((0, noisy_all.size(1)))
    return all_gather_enhanced(local, noisy_all.size(0), group)
