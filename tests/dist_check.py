"""Multi-GPU parity check (run under torchrun, one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_check.py

R ranks x (B/R samples each) with row-sharded tables must reproduce the reference's single-process step on the
whole batch: golden y_pred / gradients / post-Adam parameters of tests/golden/en_small.npz (and enp_small).
Exits non-zero on mismatch; rank 0 prints a summary."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from golden_util import Golden, subsample  # noqa: E402
from oracle import satrans_oracle as O  # noqa: E402
from satrans_b200 import DenseFeat, SATrans, SparseFeat  # noqa: E402
from satrans_b200.dist import DistContext  # noqa: E402
from test_oracle_golden import assert_adam_close  # noqa: E402


def build(g, rank, world, dev):
    m = g.meta
    cols = ([SparseFeat(n, v, 32) for n, v in zip(m['sparse_names'], m['vocab_sizes'])] +
            [DenseFeat(n, 1) for n in m['dense_names']])
    model = SATrans(cols, cols, [m['domain_col']], [m['num_domains']], att_layer_num=0,
                    domain_att_layer_num=m['domain_att_layer_num'], att_head_num=m['att_head_num'],
                    meta_mode=m['meta_mode'], seed='1021', device=dev, flag=m['flag'], table_shard=(rank, world))
    full = g.full_state_dict()
    dense_sd = {k: v for k, v in full.items() if not k.startswith('embedding_dict.')}
    model.load_state_dict(dense_sd, strict=False)
    table = torch.cat([full['embedding_dict.%s.weight' % n] for n in m['sparse_names']], dim=0)
    model.load_full_tables(table)
    model.precision = os.environ.get('SAT_PRECISION', 'tf32x3')    # package default unless overridden
    return model, table


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    dist.init_process_group('nccl', device_id=torch.device(dev))
    ctx = DistContext()
    ok = True
    for name in ('en_small', 'enp_small'):
        g = Golden(name)
        model, table0 = build(g, rank, world, dev)
        eng = model.engine
        eng.attach_dist(ctx)
        B = g.B
        per = B // world
        eng.prepare_training(per + B % world)
        lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else B
        l2 = g.cfg.l2_reg_embedding
        # ---- forward + gradients on the sharded batch ------------------------------------------------------
        X, y = g.X()[lo:hi].cuda().contiguous(), g.y()[lo:hi].cuda().contiguous()
        yp, w = eng.loss_and_grads(X, y, training=False, l2_emb=l2)
        torch.cuda.synchronize()
        err_y = (yp.cpu() - g.arr('y_pred')[lo:hi, 0]).abs().max().item()
        bce, reg = float(w['scalars'][0]), float(w['scalars'][1])
        e_b = abs(bce - float(g.z['bce'])) / abs(float(g.z['bce']))
        e_r = abs(reg - float(g.z['reg'])) / abs(float(g.z['reg']))
        names = {id(p): n for n, p in model.named_parameters()}
        flat = w['dgrad'].cpu()
        worst = 0.0
        for p, off in model._flat_slots():
            k = names[id(p)]
            ref = g.arr('grad/' + k)
            got = flat[off:off + p.numel()].view(p.shape)[subsample(k)]
            worst = max(worst, (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-30)
                        if ref.abs().max().item() > 1e-8 else 0.0)
        # table gradients of the rows this rank owns
        n = int(w['scalars'][2].item())
        rows_local = w['uniq_rows'][:n].long().cpu()
        rg = w['row_grad'][:n * 32].view(n, 32).cpu()
        ref_full = torch.cat([g.arr('grad/embedding_dict.%s.weight' % f) for f in g.cfg.sparse_names], dim=0)
        mine = 2.0 * l2 * table0[rank::world]
        mine[rows_local] += rg
        e_t = (mine - ref_full[rank::world]).abs().max().item() / ref_full.abs().max().item()
        good = err_y <= 2e-6 and e_b <= 1e-5 and e_r <= 1e-5 and worst <= 1e-4 and e_t <= 1e-4
        ok &= good
        print('[rank %d] %s: y_pred err %.2e  bce rel %.1e  reg rel %.1e  dense-grad rel %.1e  table-grad rel %.1e  %s'
              % (rank, name, err_y, e_b, e_r, worst, e_t, 'OK' if good else 'FAIL'), flush=True)
        # ---- three optimiser steps == the reference's three dense-Adam steps on the whole batches ----------
        model.load_full_tables(table0)
        model.load_state_dict({k: v for k, v in g.full_state_dict().items() if not k.startswith('embedding_dict.')},
                              strict=False)
        eng.reset_optimizer()
        lr = float(g.z['lr'])
        for step in range(1, 4):
            Xs, ys = g.X(step - 1)[lo:hi].cuda().contiguous(), g.y(step - 1)[lo:hi].cuda().contiguous()
            eng.train_step(Xs, ys, lr=lr, l2_emb=l2, dense_mode=True, training=False)
        torch.cuda.synchronize()
        full_now = model.gather_tables().cpu()
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        try:
            off = 0
            for f, v in zip(g.cfg.sparse_names, g.cfg.vocab_sizes):
                assert_adam_close(full_now[off:off + v], g.arr('step3/embedding_dict.%s.weight' % f), lr, 3, f)
                off += v
            for k in O.trainable_keys(g.cfg):
                if not k.startswith('embedding_dict.'):
                    assert_adam_close(sd[k][subsample(k)], g.arr('step3/' + k), lr, 3, k)
            print('[rank %d] %s: 3 sharded Adam steps match the reference  OK' % (rank, name), flush=True)
        except AssertionError as e:
            ok = False
            print('[rank %d] %s: Adam steps FAIL %s' % (rank, name, e), flush=True)
        # ---- deferred table sweep, row-sharded: 6 train-mode steps with the every-row update batched over epochs of 4 steps
        #      == the same steps with a sweep after each, bit for bit (tables of every shard, dense parameters) --------------
        runs = []
        for defer in (1, 4):
            model.load_full_tables(table0)
            model.load_state_dict({k: v for k, v in g.full_state_dict().items() if not k.startswith('embedding_dict.')},
                                  strict=False)
            eng.reset_optimizer()
            for step in range(6):
                Xs, ys = g.X(step % 3)[lo:hi].cuda().contiguous(), g.y(step % 3)[lo:hi].cuda().contiguous()
                eng.train_step(Xs, ys, lr=lr, l2_emb=l2, dense_mode=True, training=True, dropout_p=0.1, seed=99 + step,
                               defer=defer)
            runs.append((model.gather_tables().cpu(), model._dense.detach().cpu().clone()))     # (gather_tables syncs)
        same = torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
        ok &= same
        print('[rank %d] %s: deferred sweep (sharded) bit-identical  %s' % (rank, name, 'OK' if same else 'FAIL'), flush=True)
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    dist.destroy_process_group()
    if int(t.item()) != 0:
        sys.exit(1)
    if rank == 0:
        print('dist_check OK (world=%d)' % world)


if __name__ == '__main__':
    main()
