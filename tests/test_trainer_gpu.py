"""Trainer.sample / Trainer.eval_sample (reference: trainers/train_prior.py:645-701, trainers/base_trainer.py:446-487):
the package-level sampling wrappers on top of generate_samples_vada_2prior."""
import json
import os

import pytest
import torch

from tests.synth import synth_state_dict

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _trainer(steps=6):
    from lion_b200.config import default_prior_cfg
    from lion_b200.trainers.train_prior import Trainer
    cfg = default_prior_cfg(num_steps=steps)
    tr = Trainer(cfg)
    shp = lambda m: {k: list(v.shape) for k, v in m.state_dict().items()}
    tr.dae[0].load_state_dict(synth_state_dict(shp(tr.dae[0]), 14))
    tr.dae[1].load_state_dict(synth_state_dict(shp(tr.dae[1]), 11))
    tr.model.decoder.load_state_dict(synth_state_dict(shp(tr.model.decoder), 13))
    return tr


def test_trainer_sample_matches_the_module_function():
    from lion_b200.trainers.train_2prior import generate_samples_vada_2prior
    tr = _trainer()
    torch.manual_seed(5)
    traj = tr.sample(num_shapes=2)
    assert traj.shape == (2, 3, 2048) and torch.isfinite(traj).all()
    torch.manual_seed(5)
    img, *_ = generate_samples_vada_2prior(tr.model.latent_shape(), tr.dae, tr.diffusion_disc, tr.model, 2, False)
    assert torch.equal(traj, img.permute(0, 2, 1).contiguous())


def test_eval_sample_iterations_and_seeding(tmp_path):
    tr = _trainer()
    out = str(tmp_path / "samples.pt")
    pcs = tr.eval_sample(num_ref=5, batch_size_test=2, output_name=out)      # 5 // 2 + 1 = 3 batches of 2
    assert pcs.shape == (6, 2048, 3) and torch.isfinite(pcs).all()
    assert torch.equal(torch.load(out), pcs)
    again = tr.eval_sample(num_ref=5, batch_size_test=2)
    assert torch.equal(again, pcs), "re-seeded per batch: the run is reproducible"
    assert not torch.equal(pcs[0:2], pcs[2:4]), "each batch has its own seed"


def test_resume_reads_reference_checkpoint_keys(tmp_path):
    tr = _trainer()
    ck = str(tmp_path / "ck.pt")
    torch.save({"dae_state_dict": tr.dae.state_dict(), "vae_state_dict": tr.model.state_dict(), "epoch": 7,
                "dae_optimizer": {"state": {}}}, ck)
    tr2 = _trainer()
    assert tr2.resume(ck) == 7
    for (k, a), (_, b) in zip(tr.dae.state_dict().items(), tr2.dae.state_dict().items()):
        assert torch.equal(a, b), k
