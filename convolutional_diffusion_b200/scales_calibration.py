"""Kernel-size calibration against a trained denoiser, score-module half on the B200 kernels.

Counterpart of `/root/reference/scripts/scales_calibration.py:33-188` (`calibrate`): for `nsamps` trajectories of a model
`eps = model(t, x, label=)`, at every step evaluate the analytic score for every candidate kernel size on the model's
current x, and record the size whose score is closest to the model's (-eps/sqrt(beta)) by cosine similarity (or L2).
Differences from the reference, none of which change the result:
  * the model is passed in as a callable (the reference torch.load()s a checkpoint that is not shipped);
  * ONE score module serves all kernel sizes (`module(t, x, k=k)`), so the bank is uploaded once instead of once per size;
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
              device="cuda", precision="auto", generator=None):
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
    nk = len(kernelsizes)
    kcosine = torch.zeros(nk, nsteps, device=device)
    kdists = torch.zeros(nk, nsteps, device=device)
    k_optimals = torch.zeros(nsamps, nsteps, device=device)
    ks = torch.as_tensor(list(kernelsizes), device=device, dtype=torch.float32)
    for s in range(nsamps):
        label = torch.randint(0, nlabels, (1,), generator=generator) if conditional else None
        x = torch.randn((1, in_channels, image_size, image_size), generator=generator).to(device)
        for i in range(nsteps, 0, -1):
            t = i * torch.ones(1, device=device) / nsteps
            beta_t = schedule(t)
            eps = model(t, x, label=label if conditional else None)
            alpha_t = 1 - beta_t
            beta_prev = schedule(t - 1 / nsteps)
            alpha_prev = 1 - beta_prev
            k_estims = [mod(t, x, label=label, device=device, k=int(k)) for k in kernelsizes]
            ratio = torch.sqrt(alpha_prev / alpha_t)[:, None, None, None]
            x = x * ratio + (torch.sqrt(beta_prev[:, None, None, None]) - ratio * torch.sqrt(beta_t[:, None, None, None])) * eps
            corrected = -eps / (beta_t ** 0.5)
            for j, ke in enumerate(k_estims):
                kdists[j, i - 1] = torch.sqrt(torch.sum((corrected - ke) ** 2))
                kcosine[j, i - 1] = torch.sum(corrected * ke) / (torch.sqrt(torch.sum(corrected ** 2)) * torch.sqrt(torch.sum(ke ** 2)))
            k_optimals[s, i - 1] = ks[torch.argmin(kdists[:, i - 1])] if eval_mode == "l2_dist" else ks[torch.argmax(kcosine[:, i - 1])]
    return {"k_optimals": k_optimals.cpu(),
            "median": torch.median(k_optimals, dim=0).values.type(torch.int).cpu(),
            "mode": torch.mode(k_optimals, dim=0).values.type(torch.int).cpu()}
