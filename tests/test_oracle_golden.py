"""Pin oracle/ against outputs of the reference itself (tests/golden/, made by make_golden.py)."""
import numpy as np
import pytest
import torch

from b200 import synth
from oracle import head_oracle, rank_oracle
from oracle.swin_oracle import SwinSpec, param_shapes, swin_forward, swin_flops_per_image


@pytest.fixture(scope='module')
def swin_run(golden_dir):
    g = np.load(golden_dir / 'swin_t_arcface_b2.npz')
    spec = SwinSpec()
    sd = synth.synth_state_dict(param_shapes(spec), seed=123)
    sd = {k: v.clone().requires_grad_(not k.endswith('_mask')) for k, v in sd.items()}
    w_arc = synth.synth_tensor('add_margin.weight', (1000, 512), seed=123).requires_grad_(True)
    img = synth.synth_images(2, seed=123)
    label = synth.synth_labels(2, 1000, seed=123)
    out = head_oracle.metric_learning_forward(lambda x: swin_forward(sd, x, spec), w_arc, img, label)
    out['loss'].backward()
    return g, sd, w_arc, label, out


def test_state_dict_template():
    shapes = param_shapes(SwinSpec())
    assert len(shapes) == 168
    n_train = sum(int(np.prod(s)) for k, s in shapes.items() if not k.endswith('_mask'))
    assert n_train == 27_874_316
    assert abs(swin_flops_per_image() / 1e9 - 8.98) < 0.01


def test_swin_embeddings_match_reference(swin_run):
    g, sd, w_arc, label, out = swin_run
    assert np.array_equal(g['label'], label.numpy())
    np.testing.assert_allclose(out['emb'].detach().numpy(), g['emb'], rtol=0, atol=2e-5)
    np.testing.assert_allclose(g['emb_eval'], g['emb'], rtol=0, atol=0)   # no dropout anywhere


def test_arcface_loss_and_logits_match_reference(swin_run):
    g, sd, w_arc, label, out = swin_run
    assert abs(out['loss'].item() - float(g['loss'])) < 1e-4
    np.testing.assert_allclose(out['logits'].detach()[:, :32].numpy(), g['logits_head'], atol=2e-4)
    np.testing.assert_allclose(out['logits'].detach()[torch.arange(2), label].numpy(), g['logits_label'], atol=2e-4)


def test_gradients_match_reference(swin_run):
    g, sd, w_arc, label, out = swin_run
    names = [str(n) for n in g['grad_names']]
    ours = {}
    for n in names:
        if n == 'add_margin.weight':
            ours[n] = w_arc.grad
        else:
            ours[n] = sd[n[len('module.'):]].grad
    norms = np.array([ours[n].double().norm().item() for n in names])
    np.testing.assert_allclose(norms, g['grad_norms'], rtol=2e-3, atol=1e-7)
    for key in g.files:
        if key.startswith('grad/'):
            ref = g[key]
            got = ours[key[len('grad/'):]].numpy()
            assert np.abs(got - ref).max() <= 2e-3 * np.abs(ref).max() + 1e-7, key
    np.testing.assert_allclose(w_arc.grad[label].numpy(), g['grad_arc_rows'], rtol=1e-3, atol=1e-6)


def test_sgd_two_steps_match_reference(swin_run):
    g, sd, w_arc, label, out = swin_run
    names = [str(n) for n in g['after2_names']]
    sums = []
    for n in names:
        if n == 'add_margin.weight':
            p, lr, wd = w_arc.detach().clone(), 1e-2, 1e-4
            grad = w_arc.grad
        else:
            t = sd[n[len('module.'):]]
            p, lr, wd = t.detach().clone(), 5e-3, 0.0       # no 'fc' in any Swin key -> group 1
            grad = t.grad if t.grad is not None else None
        if grad is not None:
            bufs = head_oracle.sgd_momentum_step([p], [grad], [None], lr, 0.9, wd)
            head_oracle.sgd_momentum_step([p], [grad], bufs, lr, 0.9, wd)
        if n == 'module.mlp_head.1.bias':
            np.testing.assert_allclose(p.numpy(), g['after2_head_bias'], rtol=1e-5, atol=1e-7)
        if n == 'add_margin.weight':
            np.testing.assert_allclose(p[label].numpy(), g['after2_arc_rows'], rtol=1e-5, atol=1e-7)
        sums.append(p.double().sum().item())
    np.testing.assert_allclose(np.array(sums), g['after2_sum'], rtol=1e-4, atol=1e-3)


def test_margin_heads_small(golden_dir):
    g = np.load(golden_dir / 'heads_small.npz')
    e = torch.tensor(g['emb'], requires_grad=True)
    w = torch.tensor(g['weight'], requires_grad=True)
    lab = torch.tensor(g['label'])
    np.testing.assert_allclose(head_oracle.cosface_logits(e, w, lab).detach().numpy(), g['cosface'], atol=1e-4)
    la = head_oracle.arcface_logits(e, w, lab)
    np.testing.assert_allclose(la.detach().numpy(), g['arcface'], atol=1e-4)
    assert abs(head_oracle.focal_loss(la, lab, 2.0).item() - float(g['focal_g2'])) < 1e-4
    l0 = head_oracle.focal_loss(la, lab, 0.0)
    assert abs(l0.item() - float(g['focal_g0'])) < 1e-4
    l0.backward()
    np.testing.assert_allclose(e.grad.numpy(), g['demb'], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(w.grad.numpy(), g['dweight'], rtol=1e-3, atol=1e-5)


def test_recall_loop_matches_reference_controller(golden_dir):
    g = np.load(golden_dir / 'recall_loop.npz')
    for tag in 'abc':
        n_id, per, sigma, seed = g[f'{tag}_spec']
        emb, classes = synth.synth_embeddings(int(n_id), int(per), sigma=float(sigma), seed=int(seed))
        if tag == 'b':
            keep = torch.ones(len(classes), dtype=torch.bool)
            keep[int(n_id):int(n_id) + 20] = False
            emb, classes = emb[keep], classes[keep]
        assert emb.shape[0] == int(g[f'{tag}_n'])
        # verbatim restatement of the loop
        got = rank_oracle.recall_at_k_loop(emb, classes, (10, 100))
        assert got['Recall@K=10'] == pytest.approx(g[f'{tag}_recall'][0], abs=1e-12)
        assert got['Recall@K=100'] == pytest.approx(g[f'{tag}_recall'][1], abs=1e-12)
        # the deterministic top-k specification gives the same Recall@K
        idx, _ = rank_oracle.topk_spec(emb.numpy(), emb.numpy(), 100, exclude_self_offset=0)
        spec = rank_oracle.recall_from_topk(idx, classes.numpy(), classes.numpy(), (10, 100), exclude_self_offset=0)
        assert spec['Recall@K=10'] == pytest.approx(g[f'{tag}_recall'][0], abs=1e-12)
        assert spec['Recall@K=100'] == pytest.approx(g[f'{tag}_recall'][1], abs=1e-12)


def test_restore_dataset_order():
    emb = torch.arange(12.).reshape(6, 2)
    perm = torch.tensor([4, 1, 5, 0, 3, 2])
    outs = [{'emb': emb[perm[:3]], 'label': perm[:3] * 10, 'index': perm[:3]},
            {'emb': emb[perm[3:]], 'label': perm[3:] * 10, 'index': perm[3:]}]
    e, c = rank_oracle.restore_dataset_order(outs)
    assert torch.equal(e, emb) and torch.equal(c, torch.arange(6) * 10)
