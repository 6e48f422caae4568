"""Trainer behind the reference API (train/trainer.py:15-200): same constructor, attributes, `fit()` / `val()`
control flow, early stopping, best-model checkpointing and returned dict — with the per-batch body
(trainer.py:128-148: forward, loss, backward, optimizer.step, zero_grad) replaced by two kernel launches:
hsk_mf_train_fused (gather + score + loss + gradient scatter in one pass) and hsk_adamw_dense.

Differences from the reference, all deliberate:
  * no nn.DataParallel wrap (trainer.py:38-40): one process per GPU; when torch.distributed is initialised with more than
    one rank (torchrun), `Trainer(...)` returns a `ShardedTrainer` (hassaku_b200/train/sharded_trainer.py) unless
    conf['multi_gpu'] == 'off'
  * the three `.item()` syncs per step (trainer.py:141-143) become one device-side fp64 accumulator read once per
    epoch; the reported `epoch_train_*` values are the same quantities
  * device 'cpu' is rejected loudly — there is no CPU path
  * models / losses outside this path (ACF, ProtoMF, DeepMF, ECF ...; any loss without a fused kernel): after
    `hassaku_b200.install()` the reference's own Trainer is still reachable and `Trainer(...)` hands such a model to it
    unchanged, so the rest of `AlgorithmsEnum` keeps training as before
"""
import logging

import torch
from torch.utils import data
from tqdm import trange, tqdm

from hassaku_b200 import _C
from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
from hassaku_b200.eval.eval import evaluate_recommender_algorithm, FullEvaluator
from hassaku_b200.train.optim import DenseAdagrad, DenseAdam
from hassaku_b200.train.rec_losses import RecommenderSystemLoss
from hassaku_b200.train.trainer_step import FusedMFTrainStep


def _reference_trainer():
    """The reference's own `train.trainer.Trainer`, if `hassaku_b200.install()` replaced it (else None)."""
    import hassaku_b200
    for obj, name, val in hassaku_b200._ORIGINALS:
        if name == 'Trainer' and getattr(obj, '__name__', '') == 'train.trainer':
            return val
    return None


class Trainer:
    """Same surface as the reference `Trainer` (train/trainer.py:15-200): constructor arguments, the attributes
    `best_value / best_metrics / best_epoch / pointer_to_model / optimizer`, `fit()` and `val()`."""

    _OPTIMIZERS = {'adamw': True, 'adam': False, 'adagrad': None}  # name -> decoupled weight decay? (trainer.py:48-53)

    def __new__(cls, model=None, train_loader=None, val_loader=None, rec_loss=None, conf=None):
        fused = isinstance(model, SGDMatrixFactorization) and getattr(rec_loss, 'loss_kind', None) in _C.LOSS_KINDS
        if cls is Trainer and not fused:
            ref = _reference_trainer()
            if ref is not None:   # not an instance of cls: Python then skips cls.__init__
                logging.info(f'{type(model).__name__} / {type(rec_loss).__name__} is outside the fused MF path: '
                             f'handing it to the reference Trainer')
                return ref(model, train_loader, val_loader, rec_loss, conf)
        if cls is Trainer and fused and (conf or {}).get('multi_gpu', 'auto') != 'off':
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                # launched under torchrun with one rank per GPU: the sharded trainer is this path's nn.DataParallel
                from hassaku_b200.train.sharded_trainer import ShardedTrainer
                return ShardedTrainer(model, train_loader, val_loader, rec_loss, conf)
        return super().__new__(cls)

    def __init__(self, model: SGDMatrixFactorization, train_loader: data.DataLoader, val_loader: data.DataLoader,
                 rec_loss: RecommenderSystemLoss, conf: dict):
        """`train_loader`: any iterable of (u_idxs, i_idxs, labels) with host or device tensors (the reference
        TrainDataLoader or hassaku_b200.data.DeviceTrainLoader); `val_loader`: a FullEvalDataset loader; `conf`: the
        reference's configuration dictionary (conf/conf_parser.py keys)."""
        self.device = conf['device']
        if not str(self.device).startswith('cuda'):
            raise _C.HskError(f"device '{self.device}' requested: hassaku_b200 has no CPU path, use device: cuda")
        if not isinstance(model, SGDMatrixFactorization):
            raise TypeError('hassaku_b200.Trainer drives hassaku_b200 SGDMatrixFactorization models only')
        self.train_loader, self.val_loader = train_loader, val_loader
        self.model = self.pointer_to_model = model  # no nn.DataParallel wrap: one process per GPU
        self.model.to(self.device)
        self.rec_loss = rec_loss
        self.lr, self.wd = conf['lr'], conf['wd']
        if conf['optimizer'] not in self._OPTIMIZERS:
            raise ValueError(f"Optimizer {conf['optimizer']} not yet implemented")
        # optional extra key (default reproduces the reference): optimizer_mode: dense | lazy (row-sparse AdamW)
        if conf['optimizer'] == 'adagrad':
            self.optimizer = DenseAdagrad(self.model, lr=self.lr, weight_decay=self.wd)
        else:
            self.optimizer = DenseAdam(self.model, lr=self.lr, weight_decay=self.wd,
                                       decoupled=self._OPTIMIZERS[conf['optimizer']], mode=conf.get('optimizer_mode', 'dense'))
        self.train_step = FusedMFTrainStep(self.model, self.rec_loss, self.optimizer)

        self.n_epochs = conf['n_epochs']
        self.optimizing_metric = conf['optimizing_metric']
        self.max_patience = conf['max_patience']
        self.model_path = conf['model_path']
        self.use_wandb = conf['running_settings']['use_wandb']
        self.batch_verbose = conf['running_settings']['batch_verbose']
        self._in_tune = conf.get('_in_tune', False)

        self.best_value = self.best_metrics = self.best_epoch = None
        logging.info('Built Trainer module: ' + ', '.join(f'{k}={v}' for k, v in dict(
            n_epochs=self.n_epochs, rec_loss=self.rec_loss.name, device=self.device,
            optimizing_metric=self.optimizing_metric, model_path=self.model_path, optimizer=conf['optimizer'],
            lr=self.lr, wd=self.wd, use_wandb=self.use_wandb, batch_verbose=self.batch_verbose,
            max_patience=self.max_patience).items()))

    # ---- helpers ----
    def _report(self, log_dict: dict, epoch: int):
        post_val = getattr(self.pointer_to_model, 'post_val', None)
        if callable(post_val):
            log_dict.update(post_val(epoch))
        if self._in_tune:
            from ray.air import session
            session.report(log_dict)
        elif self.use_wandb:
            import wandb
            wandb.log(log_dict)

    def _train_one_epoch(self) -> dict:
        """trainer.py:112-150: one pass over the loader; returns the epoch-average losses."""
        self.model.train()
        n_batches = 0
        for u_idxs, i_idxs, _labels in (tqdm(self.train_loader) if self.batch_verbose else self.train_loader):
            self.train_step(u_idxs, i_idxs)
            n_batches += 1
        rec = self.train_step.pop_loss_sum() / max(n_batches, 1)  # the only host sync of the epoch
        # MF has no auxiliary loss (base_classes.py:148 returns zeros), so total == rec
        return {'epoch_train_loss': rec, 'epoch_train_rec_loss': rec, 'epoch_train_reg_loss': 0.0}

    def fit(self):
        """Validation before any update (epoch -1), then up to `n_epochs` epochs with early stopping on
        `optimizing_metric` (patience `max_patience`) and a checkpoint of every new best model — trainer.py:85-185.
        Returns `best_metrics` (includes `max_optimizing_metric` and `best_epoch`)."""
        first = self.val()
        self.best_value = first['max_optimizing_metric'] = first[self.optimizing_metric]
        self.best_epoch = first['best_epoch'] = -1
        self.best_metrics = first
        print('Init - Avg Val Value {:.3f} \n'.format(self.best_value))
        self._report(first, -1)
        self.pointer_to_model.save_model_to_path(self.model_path)

        patience = self.max_patience
        for epoch in trange(self.n_epochs):
            if patience == 0:
                print('Ran out of patience, Stopping ')
                break
            losses = self._train_one_epoch()
            print('Epoch {} - Epoch Avg Train Loss {:.4f} ({:.4f} Rec Loss + {:.4f} Reg Loss )\n'.format(
                epoch, losses['epoch_train_loss'], losses['epoch_train_rec_loss'], losses['epoch_train_reg_loss']))

            metrics = self.val()
            value = metrics[self.optimizing_metric]
            print('Epoch {} - Avg Val Value {:.4f} \n'.format(epoch, value))
            if value > self.best_value:
                self.best_value, self.best_epoch, self.best_metrics = value, epoch, metrics
                metrics['max_optimizing_metric'], metrics['best_epoch'] = value, epoch
                print('Epoch {} - New best model found (val value {:.4f}) \n'.format(epoch, value))
                self.pointer_to_model.save_model_to_path(self.model_path)
                patience = self.max_patience
            else:
                metrics['max_optimizing_metric'] = self.best_value
                patience -= 1
            self._report({**metrics, **losses}, epoch)
        return self.best_metrics

    @torch.no_grad()
    def val(self):
        """Full-rank evaluation on the validation loader (trainer.py:187-200) -> metric dict."""
        self.model.eval()
        print('Validation started')
        ds = self.val_loader.dataset
        evaluator = FullEvaluator(aggr_by_group=True, n_groups=ds.n_user_groups,
                                  user_to_user_group=ds.user_to_user_group)
        return evaluate_recommender_algorithm(self.pointer_to_model, self.val_loader, evaluator, self.device,
                                              self.batch_verbose)
