"""hassaku_b200 — the SGD matrix-factorization hot path of karapostK/hassaku (training step + full-rank evaluator) on
hand-written sm_100a CUDA kernels behind the reference's Python API.  See DESIGN.md / INTEGRATION.md.

    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.train.rec_losses import RecommenderSystemLossesEnum
    from hassaku_b200.train.trainer import Trainer
    from hassaku_b200.eval.eval import FullEvaluator, evaluate_recommender_algorithm

or, inside a checkout of the reference, `import hassaku_b200; hassaku_b200.install()` before `run_experiment.py`'s
helpers are called.
"""
__version__ = '0.1.0'
_ORIGINALS = []   # (object, attribute name, original value) for uninstall()


def install(device_loaders: bool = False):
    """Rebind the reference's own symbols (the reference repository must be importable: its top-level packages
    `algorithms`, `train`, `eval`, `data`, `experiment_helper`) to the CUDA implementations, keeping the originals
    reachable (`hassaku_b200.uninstall()` restores them).  After this an unmodified `run_experiment.py -a mf ...` with
    `device: cuda` runs the path on the kernels.  With `device_loaders=True` `data.data_utils.get_dataloader` also
    returns the device-resident loaders (otherwise the reference's host loaders keep feeding host batches).
    Returns the dict of patched (module, name) pairs."""
    import importlib

    from hassaku_b200.algorithms import sgd_alg as h_alg
    from hassaku_b200.eval import eval as h_eval
    from hassaku_b200.train import rec_losses as h_loss
    from hassaku_b200.train import trainer as h_trainer

    patched = {}

    def rebind(obj, name, new):
        if not any(o is obj and n == name for o, n, _ in _ORIGINALS):
            _ORIGINALS.append((obj, name, obj.__dict__[name]))   # the raw attribute (keeps staticmethod wrappers)
        setattr(obj, name, new)
        patched[(obj.__name__, name)] = new

    r_alg = importlib.import_module('algorithms.sgd_alg')
    r_loss = importlib.import_module('train.rec_losses')
    r_trainer = importlib.import_module('train.trainer')
    r_eval = importlib.import_module('eval.eval')
    # model factory: AlgorithmsEnum.mf.value.build_from_conf (experiment_helper.py:39)
    rebind(r_alg.SGDMatrixFactorization, 'build_from_conf', staticmethod(h_alg.SGDMatrixFactorization.build_from_conf))
    # AlgorithmsEnum.sgdbias: the bias-only baseline runs on the same kernels (embedding_dim 1, zero embedding tables)
    rebind(r_alg.SGDBaseline, 'build_from_conf', staticmethod(h_alg.SGDBaseline.build_from_conf))
    # loss factories: RecommenderSystemLossesEnum[...].value.build_from_conf (experiment_helper.py:42)
    for name in ('RecBinaryCrossEntropy', 'RecBayesianPersonalizedRankingLoss', 'RecSampledSoftmaxLoss'):
        rebind(getattr(r_loss, name), 'build_from_conf', staticmethod(getattr(h_loss, name).build_from_conf))
    # drivers
    rebind(r_trainer, 'Trainer', h_trainer.Trainer)
    for mod in (r_trainer, r_eval):
        rebind(mod, 'FullEvaluator', h_eval.FullEvaluator)
        rebind(mod, 'evaluate_recommender_algorithm', h_eval.evaluate_recommender_algorithm)
    # sweep_test.py imports the calibration decorator from eval.eval by name
    rebind(r_eval, 'FullEvaluatorCalibrationDecorator', h_eval.FullEvaluatorCalibrationDecorator)
    try:
        r_helper = importlib.import_module('experiment_helper')
        rebind(r_helper, 'Trainer', h_trainer.Trainer)
        rebind(r_helper, 'FullEvaluator', h_eval.FullEvaluator)
        rebind(r_helper, 'evaluate_recommender_algorithm', h_eval.evaluate_recommender_algorithm)
    except ImportError:  # wandb / ray not installed: the helpers are optional
        pass
    if device_loaders:
        r_data = importlib.import_module('data.data_utils')
        from hassaku_b200.data import dataloader as h_dl, dataset as h_ds

        def get_dataloader(conf: dict, split_set: str):
            if split_set == 'train':
                ds = h_ds.TrainRecDataset(conf['dataset_path'])
                sampler = h_dl.NegativeSampler(ds, conf['neg_train'], conf['train_neg_strategy'])
                return h_dl.TrainDataLoader(sampler, ds, batch_size=conf['train_batch_size'], shuffle=True,
                                            seed=conf.get('seed', 64))
            if split_set in ('val', 'test'):
                return h_dl.EvalLoader(h_ds.FullEvalDataset(conf['dataset_path'], split_set), conf['eval_batch_size'])
            raise ValueError(f"split_set value '{split_set}' is invalid! Please choose from [train, val, test]")

        rebind(r_data, 'get_dataloader', get_dataloader)
    return patched


def uninstall():
    """Undo install()."""
    while _ORIGINALS:
        obj, name, val = _ORIGINALS.pop()
        setattr(obj, name, val)
