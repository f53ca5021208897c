"""Host-side logic on the CPU: config loader, metrics, evaluation loops' [dataloader][batch] contract, synthetic data,
and the multi-GPU gather / merge plumbing under a 2-rank gloo group (kernels replaced by the oracle there)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / 'pets-face-recognition_b200'


def test_config_loader_and_trainer_factory(tmp_path, monkeypatch):
    from utils import Config, DictWrapper, configure_trainer, get_config, get_strategy, parse_gpus
    cfg_file = tmp_path / 'cfg.py'
    cfg_file.write_text('import os\nn_epochs = 3\ndevice = "cpu"\n_hidden = 1\nk = [5, 10]\n'
                        'distributed_train = not isinstance(device, str)\ntrainer_kwargs = dict(benchmark=True)\n'
                        'def model():\n    return 1\n')
    cfg = get_config(cfg_file)
    assert isinstance(cfg, Config) and cfg.n_epochs == 3 and cfg['k'] == [5, 10]
    assert cfg.get('missing') is None and cfg.get('missing', 7) == 7          # falls through to dict.get
    assert 'os' not in cfg and '_hidden' not in cfg                           # modules / private names are dropped
    assert callable(cfg.model)
    cfg.output = 'x'                                                          # main.py writes keys back
    assert cfg['output'] == 'x'
    assert get_config(cfg_file) is not cfg                                    # singleton is reset per load
    assert parse_gpus(cfg) == 0 and get_strategy(cfg) is None
    tr = configure_trainer(cfg, None, None)
    assert tr.max_epochs == 3 and tr.device.type == 'cpu' and tr.enable_checkpointing
    assert isinstance(DictWrapper({'a': 1}), DictWrapper)
    # singleton as in the reference: Config() anywhere in the process is the configuration get_config loaded last
    live = get_config(cfg_file)
    assert Config() is live and Config().n_epochs == 3
    assert set(live) >= {'n_epochs', 'device', 'k', 'model'} and '_filled' not in set(live) and len(live) == len(list(live))
    assert {i: repr(live[i]) for i in live}['n_epochs'] == '3'                # Controller builds its hparams this way
    from utils import get_dict_wrapper, is_main_process
    side = get_dict_wrapper(cfg_file)                                         # independent of the singleton (TSV scripts)
    side.n_epochs = 9
    assert Config().n_epochs == 3 and side['n_epochs'] == 9 and side.get('nope', 1) == 1
    for var in ('RANK', 'NODE_RANK', 'LOCAL_RANK'):
        monkeypatch.delenv(var, raising=False)
    assert is_main_process()
    monkeypatch.setenv('LOCAL_RANK', '0')                                     # torchrun rank 0 IS the main process (it writes the checkpoints)
    assert is_main_process()
    monkeypatch.setenv('RANK', '1')
    assert not is_main_process()


def test_shipped_configs_keep_the_reference_keys(monkeypatch):
    monkeypatch.setenv('SYNTH_TRAIN_IDS', '6')
    monkeypatch.setenv('SYNTH_VAL_IDS', '5')
    monkeypatch.setenv('SYNTH_PAIRS', '10')
    monkeypatch.chdir(PKG)
    from utils import get_config
    for path in ('configs/dog_fe/swin_t_dog_head_synth.py', 'configs/cat_fe/swin_t_cat_head_synth.py'):
        cfg = get_config(PKG / path)
        for key in ('model', 'loss', 'optimizer', 'train_dataloader', 'val_dataloader', 'pair_generator', 'similarity_f', 'k',
                    'thrs', 'n_epochs', 'train_batch_size', 'test_batch_size', 'output', 'device', 'distributed_train', 'world_size'):
            assert key in cfg, key
        from engine import Controller
        c = Controller(cfg)
        keys = list(c.state_dict().keys())
        assert len(keys) == 169 and keys[0] == 'model_loss.add_margin.weight'
        assert 'model_loss.module.stage1.layers.0.1.attention_block.fn.fn.upper_lower_mask' in keys
        optim, sched = c.configure_optimizers()
        assert [g['lr'] for g in optim[0].param_groups] == [0.005, 0.01, 0.01]
        assert optim[0].param_groups[2]['weight_decay'] == 1e-4 and len(optim[0].param_groups[1]['params']) == 0
        name, pg = cfg.pair_generator(0)
        assert name == 'Val' and len(pg.corrected_indices) == len(pg.labels) == 20
        s = cfg.similarity_f([(torch.ones(4), torch.ones(4)), (torch.ones(4), -torch.ones(4))])
        assert torch.allclose(s, torch.tensor([1.0, 0.0]))


def test_metrics_against_sklearn():
    from sklearn.metrics import average_precision_score, roc_auc_score, roc_curve
    from engine import metrics as M
    rng = np.random.RandomState(0)
    labels = rng.randint(0, 2, 500)
    scores = np.round(rng.rand(500) * 0.6 + labels * 0.25, 2)          # rounded -> ties
    s, l = torch.tensor(scores), torch.tensor(labels)
    assert M.auroc(s, l) == pytest.approx(roc_auc_score(labels, scores), abs=1e-12)
    assert M.average_precision(s, l) == pytest.approx(average_precision_score(labels, scores), abs=1e-9)
    fpr, tpr, thr = M.roc(s, l)
    f2, t2, _ = roc_curve(labels, scores, drop_intermediate=False)
    np.testing.assert_allclose(fpr.numpy(), f2, atol=1e-12)
    np.testing.assert_allclose(tpr.numpy(), t2, atol=1e-12)


class _StubModule(torch.nn.Module):
    """CPU module with the Controller's step interface but no B200 ops."""

    def __init__(self, n_loaders=2):
        super().__init__()
        self.lin = torch.nn.Linear(4, 3)
        self.n_loaders = n_loaders
        self.seen = None

    def _dl(self, n):
        data = [{'x': torch.full((4,), float(i)), 'label': i % 3, 'index': i} for i in range(n)]
        return torch.utils.data.DataLoader(data, batch_size=4)

    def test_dataloader(self):
        return [self._dl(10), self._dl(6)][:self.n_loaders] if self.n_loaders > 1 else self._dl(10)

    val_dataloader = test_dataloader

    def test_step(self, batch, batch_idx, dataset_idx=0):
        assert not self.training and not torch.is_grad_enabled()
        return {'emb': self.lin(batch['x']), 'label': batch['label'], 'index': batch['index']}

    validation_step = test_step

    def test_epoch_end(self, outputs):
        self.seen = outputs

    validation_epoch_end = test_epoch_end


@pytest.mark.parametrize('n_loaders', [1, 2])
def test_eval_loops_hand_over_dataloader_by_batch_lists(n_loaders):
    """engine/loops/eval_loop.py:30-37: *_epoch_end always receives outputs[dataloader][batch], also for ONE loader."""
    from engine import Trainer
    m = _StubModule(n_loaders)
    m.train()
    tr = Trainer(gpus=0, max_epochs=1)
    tr.test(m)
    assert isinstance(m.seen, list) and len(m.seen) == n_loaders
    assert [len(x) for x in m.seen] == [3, 2][:n_loaders]
    assert m.seen[0][0]['emb'].shape == (4, 3) and m.training            # back to train mode afterwards
    tr.validate(m)
    assert len(m.seen) == n_loaders


def test_controller_gather_restores_dataset_order():
    from engine import Controller
    emb = torch.arange(12.).reshape(6, 2)
    perm = torch.tensor([4, 1, 5, 0, 3, 2])
    outs = [{'emb': emb[perm[:3]], 'label': perm[:3] * 10, 'index': perm[:3]}, {'emb': emb[perm[3:]], 'label': perm[3:] * 10, 'index': perm[3:]}]
    e, c, local = Controller._gather(outs)
    assert torch.equal(e, emb) and torch.equal(c, torch.arange(6) * 10) and local is None       # one process: no shard bookkeeping
    sc, lab = torch.tensor([0.9, 0.8, 0.3, 0.2]), torch.tensor([1, 1, 0, 0])
    from engine import metrics as M
    fpr, tpr, thr = M.roc(sc, lab)
    assert Controller.compute_accuracy(sc, lab, thr, fpr, 1 - tpr) == 0.75   # reference's strict '>' at the optimal threshold


def test_synthetic_data_contract():
    from data_loading import SyntheticPairs, SyntheticRecDataset
    ds = SyntheticRecDataset(5, 3, image_size=28, seed=1)
    item = ds[7]
    assert set(item) == {'x', 'label', 'index'} and item['x'].shape == (3, 28, 28) and item['index'] == 7
    assert 0.0 <= float(item['x'].min()) and float(item['x'].max()) <= 1.0
    assert torch.equal(ds[7]['x'], ds[7]['x']) and ds.get_users() == [0, 1, 2, 3, 4]
    pg = SyntheticPairs(ds, 20, 1, seed=3)
    assert len(pg) == 40 and int(pg.labels.sum()) == 20
    lab = ds.labels
    assert all((lab[a] == lab[b]) == bool(y) for (a, b), y in zip(pg.corrected_indices, pg.labels))


# ---------------------------------------------------------------------------------------------------------------------
def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [str(ROOT), str(PKG)]
    import torch.distributed as dist
    from b200 import gallery, synth
    from oracle import rank_oracle
    dist.init_process_group('gloo', rank=rank, world_size=world)
    emb, classes = synth.synth_embeddings(30, 3, sigma=3.0, seed=11)          # 90 rows
    bounds = [0, 50, 90]                                                        # ragged shards
    lo, hi = bounds[rank], bounds[rank + 1]

    def topk_loo(q, g, k, off):     # oracle stands in for the CUDA kernel: the test is about gather / offsets / reduce
        return torch.from_numpy(rank_oracle.topk_spec(q.numpy(), g.numpy(), k, exclude_self_offset=off)[0])

    def hits_fn(idx, qc, gc, ks):
        first = [next((j for j, i in enumerate(row.tolist()) if i >= 0 and gc[i] == c), 10 ** 9) for row, c in zip(idx, qc.tolist())]
        return torch.tensor([sum(f < k for f in first) for k in ks])

    r = gallery.recall_at_k_sharded(emb[lo:hi], classes[lo:hi], (5, 10), topk_fn=topk_loo, hits_fn=hits_fn)

    def topk_g(q, g, k, base):
        i, s = rank_oracle.topk_spec(q.numpy(), g.numpy(), k)
        i = np.where(i >= 0, i + base, -1).astype(np.int32)
        return torch.from_numpy(i), torch.from_numpy(s)

    def merge(scores, idx, k):
        lists, nq, kin = scores.shape
        s = scores.permute(1, 0, 2).reshape(nq, -1).numpy()
        i = idx.permute(1, 0, 2).reshape(nq, -1).numpy().astype(np.int64)
        oi = np.full((nq, k), -1, np.int32); os_ = np.full((nq, k), -np.inf)
        for r_ in range(nq):
            key = np.where(i[r_] >= 0, i[r_], 2 ** 40)
            order = np.lexsort((key, -s[r_]))[:k]
            order = order[i[r_][order] >= 0]
            oi[r_, :len(order)] = i[r_][order]; os_[r_, :len(order)] = s[r_][order]
        return torch.from_numpy(oi), torch.from_numpy(os_)

    q, _ = synth.synth_embeddings(7, 3, sigma=3.0, seed=12)                    # 21 queries, sharded 12 / 9
    qb = [0, 12, 21]
    idx, score = gallery.cosine_topk_gallery_sharded(q[qb[rank]:qb[rank + 1]], emb[lo:hi], 10, topk_fn=topk_g, merge_fn=merge)
    torch.save({'recall': r, 'idx': idx, 'score': score}, Path(tmp) / f'out{rank}.pt')
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gallery_plumbing_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    from b200 import synth
    from oracle import rank_oracle
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    outs = [torch.load(tmp_path / f'out{r}.pt', weights_only=False) for r in range(2)]
    emb, classes = synth.synth_embeddings(30, 3, sigma=3.0, seed=11)
    full = rank_oracle.recall_at_k_loop(emb, classes, (5, 10))
    assert outs[0]['recall'] == outs[1]['recall']
    for k in (5, 10):
        assert outs[0]['recall'][f'Recall@K={k}'] == pytest.approx(full[f'Recall@K={k}'], abs=1e-12)
    q, _ = synth.synth_embeddings(7, 3, sigma=3.0, seed=12)
    ref_idx, ref_score = rank_oracle.topk_spec(q.numpy(), emb.numpy(), 10)
    got = torch.cat([outs[0]['idx'], outs[1]['idx']]).numpy()
    assert np.array_equal(got, ref_idx)
    np.testing.assert_allclose(torch.cat([outs[0]['score'], outs[1]['score']]).numpy(), ref_score, atol=1e-12)


# ---------------------------------------------------------------------------------------------------------------------
def _shard_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [str(ROOT), str(PKG)]
    import torch.distributed as dist
    from torch.utils.data import DataLoader, TensorDataset
    from engine import Trainer
    from utils import is_main_process
    tr = Trainer(gpus=0, strategy='ddp')                       # CPU + gloo: the host-side DDP plumbing only
    assert (tr.world_size, tr.rank) == (world, rank) and is_main_process() == (rank == 0)
    ds = TensorDataset(torch.arange(22))
    torch.manual_seed(123)                                     # what the shipped configs do on every rank alike
    seen = {}
    for shuffle in (True, False):
        base = DataLoader(ds, batch_size=4, shuffle=shuffle, drop_last=True)
        seen[shuffle] = [[int(v) for b in tr._shard_loader(base, epoch) for v in b[0]] for epoch in (0, 1)]
    torch.save(seen, Path(tmp) / f'seen{rank}.pt')
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_train_loader_is_sharded_by_rank_world2_gloo(tmp_path):
    """PL injected a DistributedSampler (replace_sampler_ddp); the Lightning-free Trainer has to do it itself,
    or every rank trains on the same batches."""
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = (torch.load(tmp_path / f'seen{r}.pt', weights_only=False) for r in range(2))
    for shuffle in (True, False):
        for epoch in (0, 1):
            ra, rb = a[shuffle][epoch], b[shuffle][epoch]
            assert len(ra) == len(rb) == 8 and not set(ra) & set(rb)          # 11 per rank -> 2 full batches of 4, disjoint shards
        assert (a[shuffle][0] != a[shuffle][1]) == shuffle                      # set_epoch reshuffles; sequential order is fixed
    assert a[False][0] == [0, 2, 4, 6, 8, 10, 12, 14]


# ---------------------------------------------------------------------------------------------------------------------
def _eval_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [str(ROOT), str(PKG)]
    import torch.distributed as dist
    from torch.utils.data import DataLoader, Dataset
    from b200 import gallery, synth
    from engine import Controller, Trainer
    from oracle import rank_oracle
    tr = Trainer(gpus=0, strategy='ddp')
    emb, classes = synth.synth_embeddings(9, 3, sigma=3.0, seed=21)           # 27 rows: ragged over 2 ranks (14 / 13)

    class DS(Dataset):
        def __len__(self):
            return emb.shape[0]

        def __getitem__(self, i):
            return {'emb': emb[i], 'label': classes[i], 'index': i}
    dl = tr._shard_eval_loader(DataLoader(DS(), batch_size=4, shuffle=False))
    outs = [b for b in dl]
    mine = torch.cat([b['index'] for b in outs])
    assert mine.tolist() == list(range(rank, 27, world))                       # strided shard, no padding / duplicates
    e_all, c_all, local = Controller._gather(outs)
    assert torch.equal(e_all, emb) and torch.equal(c_all, classes) and torch.equal(local, mine)     # whole set, dataset order, own rows

    # recall_at_k_rows with the oracle standing in for the kernels (host logic: reorder / offsets / all-reduce)
    def topk(q, g, k, exclude_self_offset=None, **_):
        return torch.from_numpy(rank_oracle.topk_spec(q.numpy(), g.numpy(), k, exclude_self_offset=exclude_self_offset)[0]), None

    def hits(idx, qc, gc, ks):
        first = [next((j for j, i in enumerate(row.tolist()) if i >= 0 and gc[i] == c), 10 ** 9) for row, c in zip(idx, qc.tolist())]
        return torch.tensor([sum(f < k for f in first) for k in ks])
    gallery.cosine_topk, gallery.recall_hits = topk, hits
    r = gallery.recall_at_k_rows(e_all, c_all, local, (2, 5))
    torch.save(r, Path(tmp) / f'r{rank}.pt')
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_eval_gather_and_recall_world2_gloo(tmp_path):
    """BASELINE.json config 5 host logic: strided evaluation shards, all-gather back into dataset order, every rank ranks its own
    queries, counts all-reduced == the reference loop over the whole set."""
    import torch.multiprocessing as mp
    from b200 import synth
    from oracle import rank_oracle
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_eval_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = (torch.load(tmp_path / f'r{r}.pt', weights_only=False) for r in range(2))
    emb, classes = synth.synth_embeddings(9, 3, sigma=3.0, seed=21)
    full = rank_oracle.recall_at_k_loop(emb, classes, (2, 5))
    assert a == b
    for k in (2, 5):
        assert a[f'Recall@K={k}'] == pytest.approx(full[f'Recall@K={k}'], abs=1e-12)


def _grad_hook_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [str(ROOT), str(PKG)]
    import torch.distributed as dist
    from engine import Trainer
    tr = Trainer(gpus=0, strategy='ddp')                       # CPU + gloo: no native engine, no peer memory -> all-reduce transport
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
    frozen = net[0].bias
    frozen.requires_grad_(False)
    assert tr._allreduce_hooks(net) == [] and tr.ddp_mode == 'nccl' and tr._peer_state is None
    tr2 = Trainer(gpus=0, strategy='ddp')
    tr2._allreduce_hooks(net)                                  # a second Trainer takes the hooks over: still ONE reduction per gradient
    x = torch.full((5, 6), float(rank + 1))
    net(x).square().sum().backward()
    grads = [p.grad.clone() for p in net.parameters() if p.requires_grad]
    torch.save({'grads': grads, 'early': len(tr2._early_reduced), 'old_early': len(tr._early_reduced), 'frozen_grad': frozen.grad},
               Path(tmp) / f'g{rank}.pt')
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_gradient_hooks_world2_gloo(tmp_path):
    """Parameters outside the native engines are reduced from post-accumulate hooks (engine/trainer.py): under gloo the summed
    gradients must equal the sum of the two ranks' single-process gradients, reduced exactly once, frozen parameters untouched."""
    import torch.multiprocessing as mp
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_grad_hook_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = (torch.load(tmp_path / f'g{r}.pt', weights_only=False) for r in range(2))
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
    net[0].bias.requires_grad_(False)
    want = None
    for r in range(2):
        net.zero_grad()
        net(torch.full((5, 6), float(r + 1))).square().sum().backward()
        g = [p.grad.clone() for p in net.parameters() if p.requires_grad]
        want = g if want is None else [w + v for w, v in zip(want, g)]
    assert a['early'] == b['early'] == 3 and a['old_early'] == 0 and a['frozen_grad'] is None
    for ga, gb, w in zip(a['grads'], b['grads'], want):
        assert torch.equal(ga, gb)
        torch.testing.assert_close(ga, w, rtol=1e-6, atol=1e-7)


def test_peer_arena_layout_and_gradient_source():
    """Host arithmetic of the peer gradient exchange (b200/peer.py): aligned segment layout, optimizer-table pointers."""
    from b200 import peer
    offsets, total = peer.build_layout([(('engine', 0), 1000), ('w', 64), ('b', 1)])
    assert offsets == {('engine', 0): 0, 'w': 1024, 'b': 1088} and total == 1152
    with pytest.raises(peer.B200Error):
        peer.build_layout([('a', 3), ('a', 4)])

    class FakeExchange:
        world, total = 4, 1152

        def grad_ptr(self, off):
            return 0x1000 + 4 * off
    prm = torch.nn.Parameter(torch.zeros(3))
    src = peer.ArenaGradSource(FakeExchange(), {id(prm): 1088})
    assert (src.n_src, src.stride, src.shift) == (4, 1152, 0) and src.ptr_of(prm) == 0x1000 + 4 * 1088
    with pytest.raises(peer.B200Error):
        src.ptr_of(torch.nn.Parameter(torch.zeros(1)))
    with pytest.raises(peer.B200Error):                       # no CPU transport: gloo runs keep the all-reduce path
        peer.PeerGradExchange(1152, torch.device('cpu'))


def test_peer_exchange_addressing_ring_order_and_double_buffer(monkeypatch):
    """The push protocol of b200/peer.py without a GPU: every bucket goes into slot `rank` of EVERY rank's arena (ring order
    starting at the next rank), the two arena buffers alternate per step, finish() hands the optimizer the shift of the buffer
    that was just filled, and segments outside the layout are refused."""
    import contextlib
    from b200 import peer
    copies, barriers = [], []

    class FakeLib:
        def b200_peer_copy(self, dst, src, nbytes, stream):
            copies.append((dst, src, nbytes, stream))
            return 0

    class Ev:
        def record(self):
            pass

    class St:
        cuda_stream = 7

        def wait_event(self, ev):
            pass

        def wait_stream(self, s):
            pass

    class Dist:
        def all_reduce(self, t, group=None):
            barriers.append(len(copies))

    class Grad:                                                # the slice of a CUDA tensor push() looks at
        is_cuda, dtype = True, torch.float32

        def __init__(self, n, ptr):
            self.n, self.ptr = n, ptr

        def is_contiguous(self):
            return True

        def numel(self):
            return self.n

        def data_ptr(self):
            return self.ptr
    monkeypatch.setattr(peer, 'lib', lambda: FakeLib())
    monkeypatch.setattr(peer.torch.cuda, 'Event', Ev)
    monkeypatch.setattr(peer.torch.cuda, 'stream', lambda s: contextlib.nullcontext())
    monkeypatch.setattr(peer.torch.cuda, 'current_stream', lambda *a: St())
    world, rank, total = 4, 1, 256
    ex = object.__new__(peer.PeerGradExchange)
    ex.world, ex.rank, ex.total, ex.buf, ex.group, ex.dist, ex._flag = world, rank, total, 0, None, Dist(), None
    ex.ptrs = [0x10000000 * (r + 1) for r in range(world)]
    ex.stream = St()
    ex._own = type('P', (), {'value': ex.ptrs[rank]})()

    ex.push(Grad(128, 0xABC0), 64)
    slot = (0 * world + rank) * total + 64
    assert copies == [(ex.ptrs[r] + 4 * slot, 0xABC0, 512, 7) for r in (2, 3, 0, 1)]       # ring order, own arena last
    assert ex.finish() == 0 and ex.buf == 1 and barriers == [4]                            # the barrier follows the copies
    copies.clear()
    ex.push(Grad(256, 0xDEF0), 0)
    slot = (1 * world + rank) * total
    assert [c[0] for c in copies] == [ex.ptrs[r] + 4 * slot for r in (2, 3, 0, 1)]         # second step: the other buffer
    assert ex.finish() == world * total and ex.buf == 0
    assert ex.grad_ptr(64) == ex.ptrs[rank] + 256                                          # optimizer table: buffer 0, slot 0
    for bad_off, n in ((200, 128), (32, 16), (-64, 16)):                                   # past the end / misaligned / negative
        with pytest.raises(peer.B200Error):
            ex.push(Grad(n, 0x1000), bad_off)
    with pytest.raises(peer.B200Error):
        g = Grad(16, 0x1000)
        g.dtype = torch.float16
        ex.push(g, 0)
