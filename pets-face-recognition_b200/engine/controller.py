"""FE Controller - drop-in for the reference's engine/controller.py (a LightningModule there, a plain
nn.Module here: Lightning is not a dependency).  Same constructor, hooks, config keys and printed metric
names; the two hot pieces run on the B200 kernels:

  training_step / validation_step / test_step  -> model_loss(...)  (models/swin.py + losses/*: native plan)
  the O(N^2) Recall@K loop of test_epoch_end / _evaluate (:77-91, :143-160) -> b200.gallery.recall_at_k

Pair scoring (similarity_f over pair_generator.corrected_indices, :60-68): when the config marks its similarity_f as the
reference's (cos + 1) / 2, one b200_pair_similarity launch over the index list, and the ROC-type metrics (engine/metrics.py)
run on the device the scores live on; otherwise the hook is called as in the reference.
"""
from pathlib import Path
from typing import Any, Optional

import torch

from b200 import gallery

from . import metrics as M


class Controller(torch.nn.Module):
    logger = None
    current_epoch = 0

    def __init__(self, config):
        super().__init__()
        self.config = config
        model = self.config.model()
        self.model_loss = self.config.loss(config, model)
        self.hparams = {i: repr(config[i]) for i in config} if hasattr(config, '__iter__') else {}

    def forward(self, *args, **kwargs) -> Any:
        return self.model_loss(*args, **kwargs)

    # ---- steps (engine/controller.py:27-46)
    def training_step(self, batch, batch_idx):
        x = batch['x']
        # config key `gpu_train_augmentation` (data_loading/gpu_augment.py): the reference's torchvision Compose
        # (configs/dog_fe/fe_dogs_config.py:17-26) applied to the uint8 batch on the device instead of per image in the
        # DataLoader workers; the backbone takes the uint8 result directly
        aug = self.config.get('gpu_train_augmentation') if hasattr(self.config, 'get') else None
        if aug is not None and torch.is_tensor(x) and x.is_cuda and x.dtype == torch.uint8:
            x = aug(x)
        loss = self.model_loss(x, batch['label'])
        return loss['loss']

    def validation_step(self, batch, batch_idx, dataset_idx=0) -> Optional[dict]:
        out = self.model_loss(batch['x'])
        return {'emb': out, 'label': batch['label'], 'index': batch['index']}

    def test_step(self, batch, batch_idx, dataset_idx=0) -> Optional[dict]:
        out = self.model_loss(batch['x'])
        return {'emb': out, 'label': batch['label'], 'index': batch['index']}

    def validation_epoch_end(self, outputs) -> None:
        self._evaluate(outputs)
        if self.logger is not None and hasattr(self.logger, 'log_artifacts'):
            self.logger.log_artifacts(str(self.config.output))

    # ---- shared pieces
    @staticmethod
    def _world():
        import torch.distributed as dist
        return (dist.get_world_size(), dist.get_rank()) if (dist.is_available() and dist.is_initialized()) else (1, 0)

    @classmethod
    def _gather(cls, outputs_i):
        """:51-56 - concatenate the per-batch dicts and undo the loader order.  Under DDP every rank holds the shard of the set
        its Trainer extracted (engine/trainer.py:_shard_eval_loader): the shards are all-gathered over NCCL and put back into
        dataset order, so everything below sees the whole set on every rank.  Returns (emb, classes, local) with `local` = the
        positions of this rank's own rows in the whole set (None on one process)."""
        emb = torch.cat([j['emb'] for j in outputs_i], dim=0)
        classes = torch.cat([j['label'] for j in outputs_i], dim=0)
        indices = torch.cat([j['index'] for j in outputs_i], dim=0)
        world, rank = cls._world()
        if world == 1:
            s = torch.argsort(indices)
            return emb[s], classes[s], None
        dev = emb.device
        emb_all, sizes = gallery._all_gather_rows(emb.float().contiguous())
        cls_all, _ = gallery._all_gather_rows(classes.to(dev).long().contiguous())
        idx_all, _ = gallery._all_gather_rows(indices.to(dev).long().contiguous())
        s = torch.argsort(idx_all)
        inv = torch.empty_like(s)
        inv[s] = torch.arange(s.numel(), device=dev)
        start = sum(sizes[:rank])
        return emb_all[s], cls_all[s], inv[start:start + sizes[rank]]

    def _pair_scores(self, emb, i):
        """:60-68.  A config whose similarity_f is the reference's (cos + 1) / 2 says so with the function attribute
        `b200_kind = 'cosine01'`; then the pairs are scored by one kernel over an index list and the scores (and every metric
        computed from them) stay on the device.  Any other similarity_f is called as the reference calls it."""
        name, pair_generator = self.config.pair_generator(i)
        labels = torch.as_tensor(pair_generator.labels)
        if emb.is_cuda and getattr(self.config.similarity_f, 'b200_kind', None) == 'cosine01':
            idx = torch.as_tensor(pair_generator.corrected_indices, dtype=torch.int64).reshape(-1, 2)
            scores = gallery.pair_similarity(emb.float(), idx[:, 0], idx[:, 1])
            return name, scores, labels.to(emb.device)
        scores = self.config.similarity_f([(emb[id1], emb[id2]) for id1, id2 in pair_generator.corrected_indices])
        return name, scores.detach().float().cpu(), labels

    def _recall_at_k(self, emb, classes, ks, local=None):
        """engine/controller.py:77-91: leave-one-out ranking of every embedding against all others.  similarity_f in
        every reference config is (cosine + 1) / 2, monotone in the cosine, so the fused cosine top-k ranks identically.
        Under DDP (`local` = this rank's rows of the gathered set) every rank ranks only ITS queries against the whole set and
        the hit / valid counts are all-reduced: BASELINE.json config 5."""
        ks = list(ks)
        if not ks:
            return {}
        dev = emb.device if emb.is_cuda else torch.device('cuda')
        emb, classes = emb.to(dev).float(), classes.to(dev).long()
        if local is None:
            return gallery.recall_at_k(emb, classes, ks)
        return gallery.recall_at_k_rows(emb, classes, local.to(dev), ks)

    # ---- test (engine/controller.py:48-93)
    def test_epoch_end(self, outputs) -> None:
        for i in range(len(outputs)):
            emb, classes, local = self._gather(outputs[i])
            name, scores, labels = self._pair_scores(emb, i)
            fpr, tpr, thresholds = M.roc(scores, labels)
            metrics = {'ROC AUC': M.auroc(scores, labels),
                       'Accuracy': self.compute_accuracy(scores, labels, thresholds, fpr, 1 - tpr)}
            metrics.update(self._recall_at_k(emb, classes, [10, 100], local))
            self.last_metrics = metrics
            if self._world()[1] == 0:
                print('', *[f'{name} {k}\t{v}' for k, v in metrics.items()], sep='\n')

    # ---- validation (engine/controller.py:95-203)
    def _evaluate(self, outputs) -> None:
        main = self._world()[1] == 0
        for i in range(len(outputs)):
            emb, classes, local = self._gather(outputs[i])
            name, scores, labels = self._pair_scores(emb, i)
            fpr, tpr, thresholds = M.roc(scores, labels)
            opt_thr = thresholds[torch.argmin(fpr + 1 - tpr)].item()
            tp, fp, tn, fn = M.stat_scores(scores, labels, opt_thr)
            if main:
                print(name, f'\nConf Mat thr = {opt_thr}', torch.tensor([[tn, fp], [fn, tp]]))
            metrics = {'ROC AUC': M.auroc(scores, labels), 'AveragePrecision': M.average_precision(scores, labels),
                       'Accuracy': self.compute_accuracy(scores, labels, thresholds, fpr, 1 - tpr), 'Opt thr': opt_thr}
            for thr in self.config.thrs:
                tp, fp, tn, fn = M.stat_scores(scores, labels, float(thr))
                metrics[f'Accuracy thr={thr}'] = (tp + tn) / max(1, tp + fp + tn + fn)
                metrics[f'Precision thr={thr}'] = tp / max(1, tp + fp)
                metrics[f'Recall thr={thr}'] = tp / max(1, tp + fn)
            metrics.update(self._recall_at_k(emb, classes, self.config.k, local))
            sorted_scores, perm = torch.sort(scores)
            sorted_labels = labels[perm]
            neg_scores, pos_scores = sorted_scores[sorted_labels == 0], sorted_scores[sorted_labels == 1]
            for far_thr in self.config.get('far_thr', ()):
                thr = neg_scores[-int(len(neg_scores) * far_thr)]
                if thr not in (0, 1):
                    metrics[f'TAR@FAR={far_thr}'] = M.stat_scores(scores, labels, thr.item())[0] / max(1, len(pos_scores))
                    metrics[f'TH@FAR={far_thr}'] = thr.item()
            for frr_thr in self.config.get('frr_thr', ()):
                thr = pos_scores[int(len(pos_scores) * frr_thr)]
                if thr not in (0, 1):
                    metrics[f'TRR@FRR={frr_thr}'] = M.stat_scores(scores, labels, thr.item())[2] / max(1, len(neg_scores))
                    metrics[f'TH@FRR={frr_thr}'] = thr.item()
            self.last_metrics = metrics
            if main:
                print(*[f'{name} {k}\t{v}' for k, v in metrics.items()], sep='\n')
            if self.logger is not None and main:
                self.logger.log_metrics({f'{name} {k}': v for k, v in metrics.items()}, self.current_epoch)

    @staticmethod
    def compute_accuracy(scores, labels, thresholds, fpr, fnr):
        gen_scores, imp_scores = scores[labels == 1], scores[labels == 0]
        t = thresholds[torch.argmin(fpr + fnr)]
        n_pairs = len(gen_scores) + len(imp_scores)
        n_true = len(gen_scores[gen_scores > t]) + len(imp_scores[imp_scores <= t])
        return n_true / n_pairs

    # ---- dataloader / optimizer passthroughs (engine/controller.py:230-246)
    def train_dataloader(self):
        return self.config.train_dataloader()

    def val_dataloader(self):
        return self.config.val_dataloader()

    def predict_dataloader(self):
        return self.test_dataloader()

    def test_dataloader(self):
        dl = self.config.get('test_dataloader')
        if dl is not None:
            return dl()
        return self.config.val_dataloader()

    def configure_optimizers(self):
        return self.config.optimizer(self.model_loss)
