"""`LION` demo wrapper -- mirror of the reference's models/lion.py:17-80 (`__init__`,
`load_model`, `sample`): VAE + [global prior, local prior] + a diffusers-style DDPM scheduler.

sample() follows the reference loop (x_T draw; per timestep: prior forward at t+1, scheduler
step, which draws its noise for t > 0 only; global latent -> style -> local prior -> decoder)
with the same order of RNG draws.  Each prior's loop replays one captured CUDA graph
(forward + noise draw + lion_scheduler_step + device-side step counter); the first step runs
eagerly (packs the model, sizes the arena) and so does the last one (t = 0, no noise draw).
The scheduler arithmetic is a restatement of diffusers 0.11.1 (PARITY UNPINNED, see
lion_b200/utils/ddpm_scheduler.py)."""
import importlib

import torch

from .. import _lib as L
from ..utils.ddpm_scheduler import DDPMScheduler
from ..utils.diffusion_pvd import DiffusionDiscretized
from .latent_points_ada_localprior import PVCNN2Prior as LocalPrior
from .vae_adain import Model as VAE


def import_model(model_str):
    """'models.score_sde.resnet.PriorSEDrop' -> class (reference: utils/model_helper.py import_model)."""
    p, m = model_str.rsplit('.', 1)
    if p.startswith('models.') or p.startswith('utils.') or p.startswith('trainers.'):
        p = 'lion_b200.' + p
    return getattr(importlib.import_module(p), m)


class LION(object):
    def __init__(self, cfg):
        self.vae = VAE(cfg).cuda()
        GlobalPrior = import_model(cfg.latent_pts.style_prior)
        global_prior = GlobalPrior(cfg.sde, cfg.latent_pts.style_dim, cfg).cuda()
        local_prior = LocalPrior(cfg.sde, cfg.shapelatent.latent_dim, cfg).cuda()
        self.priors = torch.nn.ModuleList([global_prior, local_prior])
        self.scheduler = DDPMScheduler(clip_sample=False, beta_start=cfg.ddpm.beta_1, beta_end=cfg.ddpm.beta_T,
                                       beta_schedule=cfg.ddpm.sched_mode, num_train_timesteps=cfg.ddpm.num_steps,
                                       variance_type=cfg.ddpm.model_var_type)
        self.diffusion = DiffusionDiscretized(None, None, cfg)
        self.use_cuda_graph = True
        self.last_gpu_launches = 0

    def load_model(self, model_path):
        ckpt = torch.load(model_path, weights_only=False)      # released ckpts carry optimizer state (SURVEY 8b hazards)
        self.priors.load_state_dict(ckpt['dae_state_dict'])
        self.vae.load_state_dict(ckpt['vae_state_dict'])                 # strict, like the reference (models/lion.py:32-35): all three sub-networks exist
        print(f'INFO finish loading from {model_path}')

    def _run_prior(self, prior, num_samples, shape, condition_input, clip_feat):
        T = self.scheduler.num_train_timesteps
        size = [num_samples] + list(shape)
        x = torch.randn(size=size, device='cuda').contiguous()
        dev = x.device
        n = x.numel()
        tables = self.scheduler.step_tables(dev)
        noise = torch.empty(size, device=dev, dtype=torch.float32)
        step = torch.zeros(1, device=dev, dtype=torch.int32)
        tfl = torch.zeros(num_samples, device=dev, dtype=torch.float32)
        lib = L.lib()

        def body(draw):
            pred = prior(x=x, t=tfl, condition_input=condition_input, clip_feat=clip_feat)
            if draw:
                torch.randn(size, device=dev, out=noise)
            L.check(lib.lion_scheduler_step(L.ptr(x), L.ptr(pred.contiguous()), L.ptr(noise), L.ptr(x), L.ptr(tables),
                                            L.ptr(step), n, L.stream()), "scheduler_step")
            L.check(lib.lion_ddpm_next_step(L.ptr(step), L.ptr(tfl), num_samples, L.stream()), "ddpm_next_step")

        with torch.cuda.device(dev):
            L.check(lib.lion_ddpm_set_step(L.ptr(step), L.ptr(tfl), num_samples, T - 1, L.stream()), "ddpm_set_step")
            body(T > 1)                                   # t = T-1, eager
            per_step = L.last_launches(dev) + 2
            graph = None
            if self.use_cuda_graph and getattr(prior, 'lion_graph_safe', True) and T > 3:
                with L.capture_graph() as graph:
                    body(True)
            for t in reversed(range(1, T - 1)):          # t = T-2 .. 1
                if graph is not None:
                    graph.replay()
                else:
                    body(True)
            if T > 1:
                body(False)                               # t = 0: the scheduler draws no noise
        self.last_gpu_launches += per_step * T
        return x

    @torch.no_grad()
    def sample(self, num_samples=10, clip_feat=None, save_img=False):
        self.scheduler.set_timesteps(self.scheduler.num_train_timesteps, device='cuda')
        latent_shape = self.vae.latent_shape()
        global_prior, local_prior = self.priors[0], self.priors[1]
        assert (not local_prior.mixed_prediction and not global_prior.mixed_prediction)
        if save_img:
            raise NotImplementedError("lion_b200: plotting (utils/vis_helper.py) is out of scope")
        self.priors.eval()
        self.last_gpu_launches = 0
        output_dict = {}
        z_global = self._run_prior(global_prior, num_samples, latent_shape[0], None, clip_feat)
        output_dict['z_global'] = z_global
        condition_input = self.vae.global2style(z_global)
        z_local = self._run_prior(local_prior, num_samples, latent_shape[1], condition_input, clip_feat)
        output_dict['z_local'] = z_local
        output = self.vae.sample(num_samples=num_samples, decomposed_eps=[z_global, z_local])
        output_dict['points'] = output
        return output_dict
