"""Multi-GPU (needs >= 2 devices; skipped on the 1-GPU box): the item-/user-sharded train step and evaluation on the real
kernels over NCCL against the single-GPU path on the union batch."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    import faulthandler
    faulthandler.dump_traceback_later(60, exit=True)   # a stalled collective must not hold the GPU box: dump the stack and exit
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
        from hassaku_b200.data.dataset import FullEvalDataset
        from hassaku_b200.data.synthetic import make_interactions
        from hassaku_b200.eval.eval import FullEvaluator, evaluate_recommender_algorithm
        from hassaku_b200.sharded import ShardedMF, partition_batch_by_user_owner
        from hassaku_b200.train.optim import DenseAdam
        from hassaku_b200.train.rec_losses import RecBayesianPersonalizedRankingLoss
        from hassaku_b200.train.trainer_step import FusedMFTrainStep
        U, I, d, B, N = 1201, 907, 402, 512, 20
        dev = torch.device('cuda', rank)
        torch.manual_seed(5)
        single = SGDMatrixFactorization(U, I, d, use_item_bias=True)
        with torch.no_grad():
            for p in single.parameters():
                p.copy_(torch.randn_like(p) * (1 / math.sqrt(d) if p.shape[-1] == d else 0.1))
        sd0 = {k: v.clone() for k, v in single.state_dict().items()}
        single.to(dev)
        lr, wd = 1e-3, 1e-4
        opt = DenseAdam(single, lr=lr, weight_decay=wd)
        step = FusedMFTrainStep(single, RecBayesianPersonalizedRankingLoss(), opt)
        smf = ShardedMF(U, I, d, use_item_bias=True, world=world, rank=rank, device=dev)
        smf.load_full_state_dict(sd0)
        rng = np.random.RandomState(9)
        for s in range(4):
            u = torch.from_numpy(rng.randint(0, U, B).astype(np.int64))
            i = torch.from_numpy(rng.randint(0, I, (B, N + 1)).astype(np.int64))
            step(u, i)
            ul, il = partition_batch_by_user_owner(u.to(dev), i.to(dev), world, rank)
            smf.step(ul, il, B, 'bpr', 0.0, lr, wd, exchange=('sparse', 'sparse', 'dense', 'dense')[s])
            l_single, l_sh = step.pop_loss_sum(), smf.pop_loss()
            assert abs(l_single - l_sh) <= 1e-5 * abs(l_single), (l_single, l_sh)
        sd = smf.full_state_dict()
        for n, p in single.state_dict().items():
            a, b = sd[n].double(), p.detach().cpu().double()
            err = float((a - b).abs().max() / b.abs().max())
            assert err < 1e-5 + 2e-3 * lr, (n, err)
        assert int(smf.status.item()) == 0
        # sharded evaluation == single-GPU evaluation (ids bit-exact: same fp32 kernel, order-independent merge)
        data = make_interactions(U, I, 40000, seed=2, n_user_groups=2)
        ds = FullEvalDataset.from_interactions(data.val, data.train, 'val', data.user_group, 2)

        class L:
            dataset, batch_size = ds, 256

        ref = evaluate_recommender_algorithm(single, L, FullEvaluator(True, 2, ds.user_to_user_group), dev)
        smf.load_full_state_dict({k: v.cpu() for k, v in single.state_dict().items()})
        got = smf.evaluate(data.val, data.train, FullEvaluator(True, 2, ds.user_to_user_group), batch_size=200)
        assert sorted(got) == sorted(ref)
        for k_, v in ref.items():
            assert abs(got[k_] - v) <= 1e-9, (k_, got[k_], v)
        if rank == 0:
            open(os.path.join(out_dir, 'ok'), 'w').write('ok')
    except BaseException:
        import traceback
        traceback.print_exc()
        os._exit(1)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
@pytest.mark.parametrize('world', [2])
def test_sharded_step_and_eval_match_single_gpu(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / 'ok').exists()


def _worker_graph(rank, world, port, out_dir):
    """The CUDA-graph replays of the dense and of the sparse (device-routed, fixed-capacity all-to-all) step equal their
    eager versions (fixed local batch shape); the peer exchange (item rows read from / gradients reduced into the owners'
    memory by the step kernel, CUDA IPC mappings) equals the sparse one, eager and captured."""
    import faulthandler
    faulthandler.dump_traceback_later(60, exit=True)
    models = []
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
        from hassaku_b200.sharded import ShardedMF
        U, I, d, B, N = 801, 507, 128, 256, 10
        dev = torch.device('cuda', rank)
        torch.manual_seed(5)
        full = SGDMatrixFactorization(U, I, d, use_item_bias=True)
        with torch.no_grad():
            for p in full.parameters():
                p.copy_(torch.randn_like(p) * (1 / math.sqrt(d) if p.shape[-1] == d else 0.1))
        for eager, graphed in (('dense', 'dense_graph'), ('sparse', 'sparse_graph'), ('sparse', 'peer'), ('peer', 'peer_graph')):
            a = ShardedMF(U, I, d, use_item_bias=True, world=world, rank=rank, device=dev)
            b = ShardedMF(U, I, d, use_item_bias=True, world=world, rank=rank, device=dev)
            models += [a, b]
            a.load_full_state_dict(full.state_dict()); b.load_full_state_dict(full.state_dict())
            rng = np.random.RandomState(rank)
            for s in range(5):
                u = torch.from_numpy((rng.randint(0, (U - rank + world - 1) // world, B) * world + rank).astype(np.int64)).to(dev)
                i = torch.from_numpy(rng.randint(0, I, (B, N + 1)).astype(np.int64)).to(dev)
                a.step(u, i, B * world, 'bpr', 0.0, 1e-3, 1e-4, exchange=eager)
                b.step(u, i, B * world, 'bpr', 0.0, 1e-3, 1e-4, exchange=graphed)
            torch.cuda.synchronize()
            assert a.t == b.t == 5
            a.check_status(); b.check_status()
            # same kernels and arithmetic; the fp32 atomics of the fused kernel land in a different order run to run
            for x, y in ((a.m, b.m), (a.v, b.v)):
                assert float((x - y).abs().max() / x.abs().max()) < 1e-5, eager
            err = float((a.arena - b.arena).abs().max() / a.arena.abs().max())
            # atomics order differs run to run; the arithmetic is identical.  Bound: the per-step tolerance of the parity
            # tests (rtol 1e-5 + the Adam-eps conditioning term 2e-3 lr for elements whose gradient nearly cancels)
            assert err < 1e-5 + 2e-3 * 1e-3, (eager, err)
            la, lb = a.pop_loss(), b.pop_loss()
            # same kernel eager / captured: fp64 accumulation of identical terms; sparse vs peer: two kernels (fp32 summation order)
            assert abs(la - lb) < (1e-9 if eager == graphed.replace('_graph', '') else 1e-6 * abs(la)), (eager, graphed, la, lb)
            assert float(a.g.abs().max()) == 0.0 and float(b.g.abs().max()) == 0.0
        if rank == 0:
            open(os.path.join(out_dir, 'ok'), 'w').write('ok')
    except BaseException:
        import traceback
        traceback.print_exc()       # visible even if the peer rank then stalls in a collective
        os._exit(1)                 # do not wait in destroy_process_group() for a peer that is inside a collective
    finally:
        for mdl in models:      # captured graphs hold NCCL work: destroy_process_group() blocks until they are dropped
            mdl.close()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
def test_graphed_steps_match_eager(tmp_path):
    mp.spawn(_worker_graph, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / 'ok').exists()


def _worker_tc_eval(rank, world, port, out_dir):
    """Item-sharded evaluation in BF16 / TF32 (cfg5's mode): every rank scores its item shard on the tensor cores; the
    merged result equals the single-GPU tensor-core evaluation (same operands, same K order -> same scores)."""
    import faulthandler
    faulthandler.dump_traceback_later(60, exit=True)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
        from hassaku_b200.data.dataset import FullEvalDataset
        from hassaku_b200.data.synthetic import make_interactions
        from hassaku_b200.eval.eval import FullEvaluator, evaluate_recommender_algorithm
        from hassaku_b200.sharded import ShardedMF
        U, I, d = 1201, 2907, 64
        dev = torch.device('cuda', rank)
        torch.manual_seed(5)
        single = SGDMatrixFactorization(U, I, d, use_user_bias=True, use_item_bias=True, use_global_bias=True)
        with torch.no_grad():
            for p in single.parameters():
                p.copy_(torch.randn_like(p) * (1 / math.sqrt(d) if p.shape[-1] == d else 0.1))
        sd0 = {k: v.clone() for k, v in single.state_dict().items()}
        single.to(dev)
        data = make_interactions(U, I, 60000, seed=2, n_user_groups=2)
        ds = FullEvalDataset.from_interactions(data.val, data.train, 'val', data.user_group, 2)

        class L:
            dataset, batch_size = ds, 256

        smf = ShardedMF(U, I, d, use_user_bias=True, use_item_bias=True, use_global_bias=True, world=world, rank=rank,
                        device=dev)
        smf.load_full_state_dict(sd0)
        for prec in ('bf16', 'tf32'):
            single.eval_precision = prec
            ref = evaluate_recommender_algorithm(single, L, FullEvaluator(True, 2, ds.user_to_user_group), dev)
            got = smf.evaluate(data.val, data.train, FullEvaluator(True, 2, ds.user_to_user_group), batch_size=200,
                               precision=prec)
            rep = smf.evaluate_replicated(data.val, data.train, FullEvaluator(True, 2, ds.user_to_user_group), batch_size=256,
                                          precision=prec)
            stm = smf.evaluate_streamed(data.val, data.train, FullEvaluator(True, 2, ds.user_to_user_group), batch_size=256,
                                        precision=prec)      # the peers' packed shards streamed over NVLink (CUDA IPC mappings)
            assert sorted(got) == sorted(ref) == sorted(rep) == sorted(stm)
            for k_, v in ref.items():
                assert abs(got[k_] - v) <= 1e-6, (prec, k_, got[k_], v)
                assert abs(rep[k_] - v) <= 1e-6, (prec, 'replicated', k_, rep[k_], v)
                assert abs(stm[k_] - v) <= 1e-6, (prec, 'streamed', k_, stm[k_], v)
        assert int(smf.status.item()) == 0
        smf.close()
        if rank == 0:
            open(os.path.join(out_dir, 'ok'), 'w').write('ok')
    except BaseException:
        import traceback
        traceback.print_exc()
        os._exit(1)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
def test_sharded_tensor_core_evaluation_matches_single_gpu(tmp_path):
    mp.spawn(_worker_tc_eval, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / 'ok').exists()


def _worker_trainer(rank, world, port, out_dir):
    """ShardedTrainer.fit (the multi-GPU Trainer.fit): trains, validates with the sharded evaluator, stops early, writes the
    reference-format checkpoint — and the checkpoint evaluates to the reported best metric on ONE GPU."""
    import faulthandler
    faulthandler.dump_traceback_later(120, exit=True)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    tr = None
    try:
        from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
        from hassaku_b200.data.dataloader import EvalLoader, NegativeSampler, TrainDataLoader
        from hassaku_b200.data.dataset import FullEvalDataset, TrainRecDataset
        from hassaku_b200.data.synthetic import make_interactions
        from hassaku_b200.eval.eval import FullEvaluator, evaluate_recommender_algorithm
        from hassaku_b200.train.rec_losses import RecBayesianPersonalizedRankingLoss
        from hassaku_b200.train.sharded_trainer import ShardedTrainer
        U, I, d = 1500, 6000, 64          # batch slots << items: the peer exchange is the one chosen
        dev = torch.device('cuda', rank)
        data = make_interactions(U, I, 60000, seed=4, n_user_groups=2)
        train_ds = TrainRecDataset.from_interactions(data.train, data.user_group, 2)
        val_ds = FullEvalDataset.from_interactions(data.val, data.train, 'val', data.user_group, 2)
        torch.manual_seed(7)
        model = SGDMatrixFactorization(U, I, d, use_item_bias=True)           # same seed on every rank = DataParallel's replica
        loader = TrainDataLoader(NegativeSampler(train_ds, 10), train_ds, batch_size=512, shuffle=True, device=dev)
        conf = {'device': 'cuda', 'lr': 1e-2, 'wd': 1e-5, 'optimizer': 'adamw', 'n_epochs': 8, 'optimizing_metric': 'ndcg@10',
                'max_patience': 3, 'model_path': os.path.join(out_dir, 'model'), 'seed': 3,
                'running_settings': {'use_wandb': False, 'batch_verbose': False}}
        tr = ShardedTrainer(model, loader, EvalLoader(val_ds, 512), RecBayesianPersonalizedRankingLoss(), conf)
        assert tr._exchange_name() == 'peer_graph'
        if world > 1:      # under an initialised process group the reference-named Trainer IS the sharded one
            from hassaku_b200.train.trainer import Trainer
            t2 = Trainer(model, loader, EvalLoader(val_ds, 512), RecBayesianPersonalizedRankingLoss(), conf)
            assert type(t2) is ShardedTrainer
            t2.close()
            assert type(Trainer(model, loader, EvalLoader(val_ds, 512), RecBayesianPersonalizedRankingLoss(), dict(conf, multi_gpu='off'))) is Trainer
        init = tr.val()['ndcg@10']
        best = tr.fit()
        assert best['ndcg@10'] > 1.5 * init and best['best_epoch'] >= 0, (init, best['ndcg@10'], best['best_epoch'])
        # every rank reports the same dict
        t = torch.tensor([best['ndcg@10']], dtype=torch.float64, device=dev)
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert float(lo) == float(hi)
        dist.barrier()
        if rank == 0:
            # the checkpoint is a reference-format state_dict: a fresh single-GPU model loads it and scores the same
            single = SGDMatrixFactorization(U, I, d, use_item_bias=True)
            single.load_model_from_path(conf['model_path'])
            single.to(dev)
            ref = evaluate_recommender_algorithm(single, EvalLoader(val_ds, 512),
                                                 FullEvaluator(True, 2, val_ds.user_to_user_group), dev)
            assert abs(ref['ndcg@10'] - best['ndcg@10']) <= 1e-9, (ref['ndcg@10'], best['ndcg@10'])
            open(os.path.join(out_dir, 'ok'), 'w').write('ok')
    except BaseException:
        import traceback
        traceback.print_exc()
        os._exit(1)
    finally:
        if tr is not None:
            tr.close()
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [1, 2])
def test_sharded_trainer_fit(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs >= {world} GPUs')
    mp.spawn(_worker_trainer, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / 'ok').exists()
