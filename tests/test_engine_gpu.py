"""The drop-in surface end to end on a GPU: main.py's call sequence (get_config -> Controller -> configure_trainer ->
fit) and eval_fe_*'s (load_state_dict(strict=False) -> test) on the shipped synthetic Swin-T config, with the
Recall@K the Controller prints checked against the oracle's statement of the reference loop on the same embeddings."""
import os
from pathlib import Path

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
PKG = Path(__file__).resolve().parents[1] / 'pets-face-recognition_b200'


@pytest.fixture(scope='module')
def run(tmp_path_factory):
    from b200 import abi
    abi.require_device()
    tmp = tmp_path_factory.mktemp('run')
    env = dict(SYNTH_TRAIN_IDS='8', SYNTH_VAL_IDS='6', SYNTH_PER_ID='4', SYNTH_BATCH='8', SYNTH_EPOCHS='2', SYNTH_WORKERS='0', SYNTH_PAIRS='30')
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        from engine import Controller
        from utils import configure_trainer, get_config
        cfg = get_config(PKG / 'configs/dog_fe/swin_t_dog_head_synth.py')
        controller = Controller(cfg)
        trainer = configure_trainer(cfg, None, tmp / 'ckpt')
        w0 = controller.model_loss.module.mlp_head[1].weight.detach().clone()
        trainer.fit(controller)
        yield cfg, controller, trainer, tmp, w0
    finally:
        os.chdir(cwd)
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_fit_trains_validates_and_checkpoints(run):
    cfg, controller, trainer, tmp, w0 = run
    assert trainer.global_step == 2 * 4                      # 32 images / batch 8, drop_last, 2 epochs
    w1 = controller.model_loss.module.mlp_head[1].weight.detach().cpu()
    assert torch.isfinite(w1).all() and not torch.equal(w1, w0)
    m = controller.last_metrics                              # from validation_epoch_end -> _evaluate
    for key in ('ROC AUC', 'AveragePrecision', 'Accuracy', 'Opt thr', 'Recall@K=5', 'Recall@K=10', 'Recall@K=100'):
        assert key in m, key
    assert 0.0 <= m['Recall@K=5'] <= m['Recall@K=10'] <= m['Recall@K=100'] <= 1.0
    ckpts = sorted((tmp / 'ckpt').glob('epoch=*.ckpt'))
    assert len(ckpts) == 2 and (tmp / 'ckpt' / 'trainer_state.pt').exists()
    sd = torch.load(ckpts[-1])
    assert len(sd) == 169 and 'model_loss.add_margin.weight' in sd


def test_eval_entry_sequence_and_recall_parity(run):
    cfg, controller, trainer, tmp, _ = run
    from engine import Controller
    from oracle import rank_oracle
    fresh = Controller(cfg)
    sd = torch.load(sorted((tmp / 'ckpt').glob('epoch=*.ckpt'))[-1])
    sd.pop('model_loss.add_margin.weight')                   # released checkpoints ship without the ArcFace weight (download_models.py:8)
    missing = fresh.load_state_dict(sd, strict=False)
    assert missing.missing_keys == ['model_loss.add_margin.weight'] and not missing.unexpected_keys
    metrics = trainer.test(fresh)
    assert set(metrics) >= {'ROC AUC', 'Accuracy', 'Recall@K=10', 'Recall@K=100'}
    outs = trainer.predict(fresh)                            # [dataloader][batch]
    emb = torch.cat([o['emb'] for o in outs[0]]).float().cpu()
    classes = torch.cat([o['label'] for o in outs[0]]).cpu()
    assert emb.shape == (24, 512)
    ref = rank_oracle.recall_at_k_loop(emb, classes, (10, 100))
    assert metrics['Recall@K=10'] == pytest.approx(ref['Recall@K=10'], abs=1e-12)
    assert metrics['Recall@K=100'] == pytest.approx(ref['Recall@K=100'], abs=1e-12)


def test_train_batches_host_path_matches_device_steps():
    """engine.Trainer.train_batches (the e2e leg of bench.py: pinned staging, H2D one step ahead, deferred loss read) must
    report, step for step, the losses of the same steps run on device-resident batches."""
    from b200 import abi, synth
    from engine.trainer import Trainer
    from losses import SoftmaxBasedMetricLearning
    from models import swin_t
    abi.require_device()

    def build():
        model = swin_t(num_classes=512)
        sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=123)
        model.load_state_dict(sd)
        wrap = SoftmaxBasedMetricLearning(model, num_class=64, embedding_size=512, is_focal=True, arc_margin=True)
        wrap.add_margin.weight.data.copy_(synth.synth_tensor('add_margin.weight', (64, 512), seed=123))
        wrap = wrap.cuda()

        class M(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.model_loss = wrap

            def training_step(self, batch, batch_idx):
                return self.model_loss(batch['x'], batch['label'])['loss']
        m = M()
        opt = torch.optim.SGD([p for p in wrap.parameters() if p.requires_grad], 5e-3, momentum=0.9)
        return m, opt

    host = [{'x': synth.synth_images(4, seed=10 + i), 'label': synth.synth_labels(4, 64, seed=10 + i)} for i in range(4)]
    m1, o1 = build()
    t1 = Trainer(gpus=[0], max_epochs=1)
    ref = [float(t1.run_training_batch(m1, {k: v.cuda() for k, v in b.items()}, [o1]).item()) for b in host]
    m2, o2 = build()
    t2 = Trainer(gpus=[0], max_epochs=1)
    got = t2.train_batches(m2, iter(host), [o2], read_loss_every=1)
    assert len(got) == 4 and all(abs(a - b) <= 1e-5 * max(1.0, abs(b)) for a, b in zip(got, ref)), (got, ref)
    got2 = t2.train_batches(m2, iter(host), [o2], read_loss_every=2)
    assert len(got2) == 2
