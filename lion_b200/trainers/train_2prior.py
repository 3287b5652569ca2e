"""Sampling entry point of the two-prior trainer -- mirror of the reference's
trainers/train_2prior.py:49-127 (`generate_samples_vada_2prior`, same signature and 5-tuple
return).  Training (`Trainer.train_iter`, optimisers, resume, visualisation) is out of scope.
"""
from timeit import default_timer as timer

import torch

from ..utils.diffusion_continuous import DiffusionBase
from ..utils.diffusion_pvd import DiffusionDiscretized


@torch.no_grad()
def generate_samples_vada_2prior(shape, dae, diffusion, vae, num_samples, enable_autocast, ode_eps=0.00001,
                                 ode_solver_tol=1e-5, ode_sample=False, prior_var=1.0, temp=1.0, vae_temp=1.0, noise=None,
                                 need_denoise=False, ddim_step=0, clip_feat=None, cls_emb=None, ddim_skip_type='uniform',
                                 ddim_kappa=1.0):
    output = {}
    assert cls_emb is None, 'lion_b200: class-conditional sampling (cls_emb) is not part of the shipped prior configs'
    if ode_sample == 1:
        # probability-flow ODE route (train_2prior.py:64-80): both priors through the adaptive host-side RK45 solver
        assert isinstance(diffusion, DiffusionBase), 'ODE-based sampling requires cont. diffusion!'
        assert ode_eps is not None and ode_solver_tol is not None
        start = timer()
        condition_input, eps_list = None, []
        nfe, time_ode_solve = 0, 0.0
        for i in range(len(dae)):
            eps, nfe, time_ode_solve = diffusion.sample_model_ode(dae[i], num_samples, shape[i], ode_eps, ode_solver_tol,
                                                                  enable_autocast, temp, noise, condition_input=condition_input,
                                                                  clip_feat=clip_feat)
            condition_input = eps
            eps_list.append(eps)
            output['sampled_eps'] = eps
        eps = vae.compose_eps(eps_list)
        return _finish(output, eps, vae, num_samples, cls_emb, start, nfe, time_ode_solve)
    assert isinstance(diffusion, DiffusionDiscretized), 'Regular sampling requires disc. diffusion!'
    assert noise is None, 'Noise is not used in ancestral sampling.'
    nfe = diffusion._diffusion_steps
    time_ode_solve = 999.999
    start = timer()
    condition_input = None
    all_eps = []
    eps_list = None
    for i in range(len(dae)):
        if ddim_step > 0:
            eps, eps_list = diffusion.run_ddim(dae[i], num_samples, shape[i], temp, enable_autocast, is_image=False,
                                               prior_var=prior_var, ddim_step=ddim_step, condition_input=condition_input,
                                               clip_feat=clip_feat, skip_type=ddim_skip_type, kappa=ddim_kappa)
        else:
            eps, eps_list = diffusion.run_denoising_diffusion(dae[i], num_samples, shape[i], temp, enable_autocast,
                                                              is_image=False, prior_var=prior_var,
                                                              condition_input=condition_input, clip_feat=clip_feat)
        condition_input = eps
        if i == 0:
            condition_input = vae.global2style(condition_input)
        all_eps.append(eps)
        output['sampled_eps'] = eps
    eps = vae.compose_eps(all_eps)
    output['eps_list'] = eps_list
    return _finish(output, eps, vae, num_samples, cls_emb, start, nfe, time_ode_solve)


def _finish(output, eps, vae, num_samples, cls_emb, start, nfe, time_ode_solve):
    """the shared tail of both routes (train_2prior.py:112-127): statistics, decoder, timing tensors"""
    output['print/sample_mean_global'] = eps.view(num_samples, -1).mean(-1).mean()
    output['print/sample_var_global'] = eps.view(num_samples, -1).var(-1).mean()
    decomposed_eps = vae.decompose_eps(eps)
    image = vae.sample(num_samples=num_samples, decomposed_eps=decomposed_eps, cls_emb=cls_emb)
    end = timer()
    sampling_time = end - start
    nfe_torch = torch.tensor(nfe * 1.0, device='cuda')
    sampling_time_torch = torch.tensor(sampling_time * 1.0, device='cuda')
    time_ode_solve_torch = torch.tensor(time_ode_solve * 1.0, device='cuda')
    return image, nfe_torch, time_ode_solve_torch, sampling_time_torch, output
