"""Kernel-size calibration against a trained denoiser, score-module half on the B200 kernels.

Counterpart of `/root/reference/scripts/scales_calibration.py:33-188` (`calibrate`): for `nsamps` trajectories of a model
`eps = model(t, x, label=)`, at every step evaluate the analytic score for every candidate kernel size on the model's
current x, and record the size whose score is closest to the model's (-eps/sqrt(beta)) by cosine similarity (or L2).
Differences from the reference, none of which change the result:
  * the model is passed in as a callable (the reference torch.load()s a checkpoint that is not shipped);
  * ONE score module serves all kernel sizes (`forward_multi_k`), so the bank is uploaded once instead of once per size,
    and all trajectories advance together: every pass over the bank serves all of them (see `samplebatch`);
  * the index convention is the reference's: the size chosen at t = i/nsteps is stored at column i-1 (:176-178),
    while ScheduledScoreMachine reads scales[i] at t = i/nsteps (idealscore.py:91,95).
"""
from __future__ import annotations

import torch

from .els_script import build_module
from .modules import cosine_noise_schedule


@torch.no_grad()
def calibrate(model, dataset, kernelsizes, scoremoduletype="bbELS", conditional=False, scorebatchsize=8, nsamps=20,
              nsteps=20, nlabels=10, eval_mode="cos", maxsamps=100000, in_channels=None, image_size=None,
              device="cuda", precision="auto", generator=None, samplebatch=None):
    """samplebatch: trajectories advanced together (default: all nsamps).  The reference runs them one after the other
    (1 600 passes over the bank for 8 sizes x 20 steps x 10 samples); they are independent, so here every pass over the
    bank serves `samplebatch` samples at once (160 passes), the candidate sizes of a step are evaluated back to back by
    `forward_multi_k`, and the comparison with the model's score stays on the device.  Initial noise and labels are drawn
    sample by sample in the reference's order, so the result does not depend on samplebatch."""
    if kernelsizes is None:
        raise ValueError("kernelsizes must be provided")
    device = torch.device(device)
    if isinstance(dataset, (tuple, list)):
        images, labels = dataset
        if maxsamps < len(images):
            dataset = (images[:maxsamps], labels[:maxsamps])
        n = len(dataset[0])
        in_channels = in_channels or int(dataset[0].shape[1])
        image_size = image_size or int(dataset[0].shape[-1])
    else:
        if maxsamps < len(dataset):
            dataset = torch.utils.data.Subset(dataset, list(range(maxsamps)))
        n = len(dataset)
    schedule = cosine_noise_schedule
    mod = build_module(scoremoduletype, dataset, n, image_size, in_channels, scorebatchsize, None, False,
                       precision=precision)
    kernelsizes = [int(k) for k in kernelsizes]
    ks = torch.as_tensor(kernelsizes, device=device, dtype=torch.float32)
    # the reference's draws, in its order: per sample first the label, then the initial noise (:128-135)
    labels0, xs0 = [], []
    for _ in range(nsamps):
        labels0.append(torch.randint(0, nlabels, (1,), generator=generator) if conditional else None)
        xs0.append(torch.randn((1, in_channels, image_size, image_size), generator=generator))
    k_optimals = torch.zeros(nsamps, nsteps, device=device)
    sb = nsamps if samplebatch is None else max(1, int(samplebatch))
    for s0 in range(0, nsamps, sb):
        rows = range(s0, min(nsamps, s0 + sb))
        x = torch.cat([xs0[s] for s in rows]).to(device)
        label = torch.cat([labels0[s] for s in rows]).to(device) if conditional else None
        B = x.shape[0]
        for i in range(nsteps, 0, -1):
            t = i * torch.ones(B, device=device) / nsteps
            beta_t = schedule(t)
            eps = model(t, x, label=label if conditional else None)
            alpha_t = 1 - beta_t
            beta_prev = schedule(t - 1 / nsteps)
            alpha_prev = 1 - beta_prev
            est = mod.forward_multi_k(t, x, kernelsizes, label=label, device=device)          # [nk, B, C, H, W]
            ratio = torch.sqrt(alpha_prev / alpha_t)[:, None, None, None]
            x = x * ratio + (torch.sqrt(beta_prev[:, None, None, None]) - ratio * torch.sqrt(beta_t[:, None, None, None])) * eps
            corrected = (-eps / (beta_t ** 0.5)[:, None, None, None])[None]                      # [1, B, C, H, W]
            kdists = torch.sqrt(torch.sum((corrected - est) ** 2, dim=(2, 3, 4)))               # [nk, B]
            kcos = torch.sum(corrected * est, dim=(2, 3, 4)) / (
                torch.sqrt(torch.sum(corrected ** 2, dim=(2, 3, 4))) * torch.sqrt(torch.sum(est ** 2, dim=(2, 3, 4))))
            best = torch.argmin(kdists, dim=0) if eval_mode == "l2_dist" else torch.argmax(kcos, dim=0)
            k_optimals[s0:s0 + B, i - 1] = ks[best]
    return {"k_optimals": k_optimals.cpu(),
            "median": torch.median(k_optimals, dim=0).values.type(torch.int).cpu(),
            "mode": torch.mode(k_optimals, dim=0).values.type(torch.int).cpu()}
