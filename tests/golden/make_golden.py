"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

/root/reference does not exist on the GPU box, so nothing at test time imports it; the vectors
written here are committed and pin oracle/ (tests/test_oracle_golden.py) and the CUDA path.

What runs unmodified from /root/reference:
  * models/swin.py                swin_t(num_classes=512)
  * losses/__init__.py, large_margin.py, losses.py    SoftmaxBasedMetricLearning(arc / cos margin)
  * engine/controller.py          Controller.test_epoch_end (the Recall@K loop), loaded by file path
                                  with pytorch_lightning / torchmetrics / matplotlib / mlflow replaced
                                  by minimal stand-ins (they are not installed; none of them takes
                                  part in the Recall@K arithmetic).
Weights/inputs come from b200.synth (seeded per key), so the fixtures hold only outputs.
"""
import contextlib
import importlib.util
import io
import os
import re
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
REF = Path('/root/reference')
sys.dont_write_bytecode = True
sys.path.insert(0, str(ROOT / 'pets-face-recognition_b200'))
sys.path.insert(0, str(ROOT))

from b200 import synth                      # noqa: E402
from oracle.swin_oracle import SwinSpec, param_shapes   # noqa: E402

OUT = Path(__file__).resolve().parent


def load_ref_module(name, rel):
    spec = importlib.util.spec_from_file_location(name, REF / rel)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def ref_models_and_losses():
    sys.path.insert(0, str(REF))
    for m in [k for k in sys.modules if k.split('.')[0] in ('models', 'losses')]:
        del sys.modules[m]
    swin = load_ref_module('ref_swin', 'models/swin.py')
    import losses as ref_losses             # /root/reference/losses (torch only)
    sys.path.remove(str(REF))
    return swin, ref_losses


def golden_swin_arcface(swin, ref_losses):
    torch.manual_seed(0)
    spec = SwinSpec()
    B, C = 2, 1000
    sd = synth.synth_state_dict(param_shapes(spec), seed=123)
    model = swin.swin_t(num_classes=512)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert list(model.state_dict().keys()) == list(param_shapes(spec).keys()), 'key order differs'
    wrap = ref_losses.SoftmaxBasedMetricLearning(model, num_class=C, embedding_size=512,
                                                 is_focal=True, arc_margin=True)
    w_arc = synth.synth_tensor('add_margin.weight', (C, 512), seed=123)
    wrap.add_margin.weight.data.copy_(w_arc)
    img = synth.synth_images(B, seed=123)
    label = synth.synth_labels(B, C, seed=123)

    wrap.eval()
    with torch.no_grad():
        emb_eval = wrap(img)
    wrap.train()
    out = wrap(img, label)
    out['loss'].backward()

    res = {'emb': out['emb'].detach().numpy(), 'emb_eval': emb_eval.numpy(),
           'loss': np.float64(out['loss'].item()),
           'logits_head': out['logits'].detach()[:, :32].numpy(),
           'logits_label': out['logits'].detach()[torch.arange(B), label].numpy(),
           'label': label.numpy()}
    names, gnorm = [], []
    for n, p in wrap.named_parameters():
        if p.grad is None:
            continue
        names.append(n)
        gnorm.append(p.grad.double().norm().item())
        if n.endswith('pos_embedding') or n in ('module.mlp_head.1.bias', 'module.stage1.patch_partition.linear.bias',
                                                 'module.stage4.layers.0.1.attention_block.fn.norm.weight',
                                                 'module.stage1.layers.0.0.mlp_block.fn.norm.bias',
                                                 'module.stage3.layers.2.1.attention_block.fn.fn.to_out.bias'):
            res['grad/' + n] = p.grad.numpy().copy()
    res['grad_names'] = np.array(names)
    res['grad_norms'] = np.array(gnorm)
    res['grad_arc_rows'] = wrap.add_margin.weight.grad[label].numpy().copy()

    # one optimizer step exactly as configs/dog_fe/fe_dogs_config.py:123-133 builds it
    params1 = [p for i, p in wrap.module.named_parameters() if 'fc' not in i]
    params2 = [p for i, p in wrap.module.named_parameters() if 'fc' in i]
    groups = [{'lr': 10 ** -2 / 2, 'params': params1}, {'lr': 10 ** -2, 'params': params2},
              {'lr': 10 ** -2, 'params': wrap.add_margin.parameters(), 'weight_decay': 1 * (10 ** -4)}]
    optim = torch.optim.SGD(groups, 0.01, momentum=0.9)
    optim.step()
    # second step with the same grads exercises the momentum buffer
    optim.step()
    res['after2_sum'] = np.array([p.detach().double().sum().item() for _, p in wrap.named_parameters()])
    res['after2_names'] = np.array([n for n, _ in wrap.named_parameters()])
    res['after2_head_bias'] = wrap.module.mlp_head[1].bias.detach().numpy().copy()
    res['after2_arc_rows'] = wrap.add_margin.weight.detach()[label].numpy().copy()
    np.savez_compressed(OUT / 'swin_t_arcface_b2.npz', **res)
    print('swin_t_arcface_b2: loss', res['loss'], 'emb[0,:4]', res['emb'][0, :4])

    # cos-margin head + CrossEntropy variant (AddMarginProduct, is_focal=False) on fixed embeddings
    torch.manual_seed(1)
    head = ref_losses.AddMarginProduct(512, 40, s=64.0, m=0.5)
    lab = torch.tensor([0, 5, 39, 7, 7, 12])
    # rows near their class centre (p not ~0, so gamma matters), one row far (margin 'else' branch)
    e = head.weight.detach()[lab] * 3.0 + 0.02 * torch.randn(6, 512)
    e[3] = -e[3] + 0.05 * torch.randn(512)
    e.requires_grad_(True)
    lg = head(e, lab)
    arc = ref_losses.ArcMarginProduct(512, 40, s=64.0, m=0.5)
    arc.weight.data.copy_(head.weight.data)
    la = arc(e, lab)
    fl = ref_losses.FocalLoss(40, gamma=2)(la, lab)
    fl0 = ref_losses.FocalLoss(40)(la, lab)
    fl0.backward()
    np.savez_compressed(OUT / 'heads_small.npz', emb=e.detach().numpy(), weight=head.weight.detach().numpy(),
                        label=lab.numpy(), cosface=lg.detach().numpy(), arcface=la.detach().numpy(),
                        focal_g2=np.float64(fl.item()), focal_g0=np.float64(fl0.item()),
                        demb=e.grad.numpy(), dweight=arc.weight.grad.numpy())
    print('heads_small: focal g0', fl0.item(), 'g2', fl.item())


def _install_stubs():
    pl = types.ModuleType('pytorch_lightning')

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *_, **__):
            pass
    pl.LightningModule = LightningModule
    loggers = types.ModuleType('pytorch_lightning.loggers')
    loggers.MLFlowLogger = type('MLFlowLogger', (), {})
    utilities = types.ModuleType('pytorch_lightning.utilities')
    ptypes = types.ModuleType('pytorch_lightning.utilities.types')
    for n in ('STEP_OUTPUT', 'EPOCH_OUTPUT', 'TRAIN_DATALOADERS', 'EVAL_DATALOADERS'):
        setattr(ptypes, n, object)
    plt_pkg = types.ModuleType('matplotlib')
    plt = types.ModuleType('matplotlib.pyplot')
    skm = types.ModuleType('sklearn.metrics')
    skm.ConfusionMatrixDisplay = type('ConfusionMatrixDisplay', (), {})
    sk = types.ModuleType('sklearn')
    tm = types.ModuleType('torchmetrics')

    class AUROC:
        def __call__(self, scores, labels):
            s, l = scores.double(), labels.bool()
            pos, neg = s[l], s[~l]
            gt = (pos[:, None] > neg[None, :]).double().sum() + 0.5 * (pos[:, None] == neg[None, :]).double().sum()
            return gt / (len(pos) * len(neg))

    class ROC:
        def __call__(self, scores, labels):
            thr = torch.unique(scores).flip(0)
            l = labels.bool()
            tpr = torch.stack([(scores[l] >= t).float().mean() for t in thr])
            fpr = torch.stack([(scores[~l] >= t).float().mean() for t in thr])
            return fpr, tpr, thr
    tm.AUROC, tm.ROC = AUROC, ROC
    for n in ('AveragePrecision', 'Recall', 'Precision', 'StatScores', 'Accuracy', 'ConfusionMatrix'):
        setattr(tm, n, type(n, (), {}))
    sys.modules.update({'pytorch_lightning': pl, 'pytorch_lightning.loggers': loggers,
                        'pytorch_lightning.utilities': utilities, 'pytorch_lightning.utilities.types': ptypes,
                        'matplotlib': plt_pkg, 'matplotlib.pyplot': plt, 'torchmetrics': tm})
    if 'sklearn.metrics' not in sys.modules:
        sys.modules.update({'sklearn': sk, 'sklearn.metrics': skm})


def golden_recall_loop():
    _install_stubs()
    ctrl_mod = load_ref_module('ref_controller', 'engine/controller.py')

    def sim(pairs):   # configs/dog_fe/fe_dogs_config.py:89-93, verbatim semantics via the config hook
        import torch.nn.functional as F
        t1 = torch.cat([i[0].unsqueeze(0) for i in pairs], dim=0)
        t2 = torch.cat([i[1].unsqueeze(0) for i in pairs], dim=0)
        return (F.cosine_similarity(t1, t2) + 1) / 2

    cases = {}
    for tag, (n_id, per, sigma, seed) in {'a': (40, 4, 3.0, 123), 'b': (60, 2, 3.5, 7), 'c': (25, 5, 4.0, 99)}.items():
        emb, classes = synth.synth_embeddings(n_id, per, sigma=sigma, seed=seed)
        if tag == 'b':      # some identities with a single image -> not counted as valid queries
            keep = torch.ones(len(classes), dtype=torch.bool)
            keep[n_id:n_id + 20] = False
            emb, classes = emb[keep], classes[keep]
        n = emb.shape[0]
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed))
        rng = np.random.RandomState(seed)
        ii = rng.randint(0, n, size=(200, 2))

        class PG:
            corrected_indices = [tuple(map(int, r)) for r in ii]
            labels = [int(classes[a] == classes[b]) for a, b in ii]
        if sum(PG.labels) == 0:
            PG.labels[0] = 1

        class Cfg(dict):
            similarity_f = staticmethod(sim)

            @staticmethod
            def model():
                return torch.nn.Identity()

            @staticmethod
            def loss(cfg, model):
                return model

            @staticmethod
            def pair_generator(i):
                return 'Val', PG
        ctrl = ctrl_mod.Controller(Cfg())
        # batches of 20 in shuffled loader order, as test_step would emit them (engine/controller.py:42-46)
        outs = [[{'emb': emb[perm[s:s + 20]], 'label': classes[perm[s:s + 20]], 'index': perm[s:s + 20]}
                 for s in range(0, n, 20)]]
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ctrl.test_epoch_end(outs)
        txt = buf.getvalue()
        r10 = float(re.search(r'Recall@K=10\t([0-9.eE+-]+)', txt).group(1))
        r100 = float(re.search(r'Recall@K=100\t([0-9.eE+-]+)', txt).group(1))
        cases[f'{tag}_spec'] = np.array([n_id, per, sigma, seed], dtype=np.float64)
        cases[f'{tag}_recall'] = np.array([r10, r100])
        cases[f'{tag}_n'] = np.int64(n)
        print('recall loop', tag, 'N', n, 'R@10', r10, 'R@100', r100)
    np.savez_compressed(OUT / 'recall_loop.npz', **cases)


if __name__ == '__main__':
    assert REF.exists(), 'the reference is only mounted in the build container'
    swin, ref_losses = ref_models_and_losses()
    golden_swin_arcface(swin, ref_losses)
    golden_recall_loop()
