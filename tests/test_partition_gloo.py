"""Host-side logic of the partitioned (multi-GPU) path, covered without GPUs: world_size 2 and 3 process groups over
gloo run the partition plan (fb_plan_partition, the same code fb_create_partitioned uses), exchange halos with
torch.distributed point-to-point calls exactly as the NCCL path does (owned boundary values out, ghost values in),
and check a distributed SpMV + dot product of the oracle's stiffness matrix against the global result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mesh, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fembrain_b200 as fb
        from oracle import pyoracle

        if isinstance(mesh, int):
            v, t, fixed, load = cases.cube_case(mesh)
        else:
            v, t, fixed = cases.golden_mesh(mesh)  # unstructured numbering: the plan is cut from a Cuthill-McKee ordering
        nV = len(v)
        plan = fb.plan_partition(nV, t, world, rank)
        b, e, l2g = plan["begin"], plan["end"], plan["l2g"]
        order, reordered = fb.partition_ordering(nV, t, world)
        assert reordered == (not isinstance(mesh, int)) or isinstance(mesh, int)
        assert np.array_equal(np.sort(order), np.arange(nV))  # a permutation, identical on every rank
        pos = np.empty(nV, np.int64)
        pos[order] = np.arange(nV)  # caller's vertex id -> position in the partition ordering
        # 1. ranges tile [0, nV) and every rank computes the same boundaries
        rng = torch.tensor([b, e])
        allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, rng)
        bounds = [int(a[0]) for a in allr] + [int(allr[-1][1])]
        assert bounds[0] == 0 and bounds[-1] == nV and all(int(allr[i][1]) == int(allr[i + 1][0]) for i in range(world - 1))
        # 2. local mesh = exactly the tets touching an owned vertex; l2g = their vertices, ascending
        touch = ((pos[t] >= b) & (pos[t] < e)).any(axis=1)
        assert np.array_equal(plan["local_tets"], np.nonzero(touch)[0])
        mine = np.unique(t[touch])
        assert np.array_equal(l2g, mine[np.argsort(pos[mine])])  # local order = ascending position in the ordering
        # 3. halo lists are mirror images across ranks (send of a to b == recv of b from a), ghosts are complete
        ghosts = l2g[(pos[l2g] < b) | (pos[l2g] >= e)]
        got = np.concatenate([plan["recv"][p] for p in plan["neighbours"]]) if plan["neighbours"] else np.zeros(0, np.int32)
        assert np.array_equal(np.sort(got), np.sort(ghosts))
        for p in plan["neighbours"]:
            mine = torch.from_numpy(plan["send"][p].astype(np.int64))
            theirs = torch.zeros(len(plan["recv"][p]), dtype=torch.int64)
            reqs = [dist.isend(mine, p), dist.irecv(theirs, p)]
            [r.wait() for r in reqs]
            assert np.array_equal(theirs.numpy(), plan["recv"][p])
        # 4. distributed SpMV + dot with halo exchange against the global product (oracle's K at a perturbed state)
        ora = pyoracle.Oracle(v, t, fixed, kind="port")
        ia, ja, _ = ora.K_csr(values=False)
        _, a = ora.force_and_matrix(cases.perturbation(v, 1.0, 1))
        rs = np.random.default_rng(3)
        xg = rs.standard_normal(3 * nV)
        yg = np.array([a[ia[i]:ia[i + 1]] @ xg[ja[ia[i]:ia[i + 1]]] for i in range(3 * nV)])
        g2l = -np.ones(nV, np.int64)
        g2l[l2g] = np.arange(len(l2g))
        xl = np.zeros(3 * len(l2g))
        own = (pos[l2g] >= b) & (pos[l2g] < e)
        for k in range(3):
            xl[3 * np.nonzero(own)[0] + k] = xg[3 * l2g[own] + k]  # owned entries only; ghosts arrive by exchange
        reqs, bufs = [], {}
        for p in plan["neighbours"]:
            sidx = g2l[plan["send"][p]]
            sb = torch.from_numpy(np.stack([xl[3 * sidx + k] for k in range(3)], axis=1).copy())
            bufs[p] = torch.zeros(len(plan["recv"][p]), 3, dtype=torch.float64)
            reqs += [dist.isend(sb, p), dist.irecv(bufs[p], p)]
        [r.wait() for r in reqs]
        for p in plan["neighbours"]:
            ridx = g2l[plan["recv"][p]]
            for k in range(3):
                xl[3 * ridx + k] = bufs[p][:, k].numpy()
        yl = np.zeros(3 * nV)
        owned_ids = order[b:e]
        for gv in owned_ids:
            for k in range(3):
                i = 3 * gv + k
                cols = ja[ia[i]:ia[i + 1]]
                lc = 3 * g2l[cols // 3] + cols % 3
                assert (g2l[cols // 3] >= 0).all()  # every column of an owned row is local (owned or ghost)
                yl[i] = a[ia[i]:ia[i + 1]] @ xl[lc]
        part = torch.tensor([float(yl @ xg)], dtype=torch.float64)  # yl is zero outside the owned rows
        dist.all_reduce(part)
        yt = torch.from_numpy(yl)
        dist.all_reduce(yt)
        assert np.allclose(yt.numpy(), yg, rtol=1e-13, atol=1e-9 * np.abs(yg).max())
        assert abs(float(part) - float(yg @ xg)) <= 1e-12 * abs(float(yg @ xg))
        out_q.put((rank, "ok"))
    except Exception as ex:  # surface the failure in the parent
        out_q.put((rank, f"{type(ex).__name__}: {ex}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mesh", [(2, 7), (3, 7), (3, "beam3")])
def test_partition_plan_and_halo_exchange_over_gloo(world, mesh):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mesh, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(r[1] == "ok" for r in res), res


def test_unstructured_numbering_is_reordered_and_the_cut_shrinks():
    """blobtree/eggshell.veg (numbering as FemBrain's polygonizer + TetGen left it): row blocks of the caller's numbering make
    nearly every vertex a ghost of every rank; the Cuthill-McKee ordering chosen on the host cuts that by more than half."""
    import fembrain_b200 as fb

    v, t, fixed = cases.golden_mesh("eggshell")
    nV = len(v)
    inc = np.bincount(t.ravel(), minlength=nV)
    for world in (2, 4, 8):
        order, reordered = fb.partition_ordering(nV, t, world)
        assert reordered and np.array_equal(np.sort(order), np.arange(nV))
        bounds = np.searchsorted(np.cumsum(inc), inc.sum() * np.arange(1, world) / world)  # the caller's numbering, same rule
        owner = np.searchsorted(bounds, np.arange(nV), side="right")
        ghosts_identity = sum(len(np.unique(t[(owner[t] == r).any(axis=1)])) - int((owner == r).sum()) for r in range(world))
        ghosts, owned = 0, 0
        for r in range(world):
            p = fb.plan_partition(nV, t, world, r)
            ghosts += len(p["l2g"]) - (p["end"] - p["begin"])
            owned += p["end"] - p["begin"]
        assert owned == nV
        assert ghosts < 0.5 * ghosts_identity, (world, ghosts, ghosts_identity)
    cube_v, cube_t, _, _ = cases.cube_case(33)
    assert not fb.partition_ordering(len(cube_v), cube_t, 2)[1]  # slabs of the structured cube are kept


def test_partition_plan_balances_incidences():
    import fembrain_b200 as fb

    v, t, fixed, _ = cases.cube_case(12)
    inc = np.bincount(t.ravel(), minlength=len(v))
    for world in (2, 4, 8):
        loads = []
        order, _ = fb.partition_ordering(len(v), t, world)
        for r in range(world):
            p = fb.plan_partition(len(v), t, world, r)
            loads.append(inc[order[p["begin"]:p["end"]]].sum())
        assert sum(loads) == inc.sum()
        assert max(loads) <= 1.25 * inc.sum() / world
    one = fb.plan_partition(len(v), t, 1, 0)
    assert one["begin"] == 0 and one["end"] == len(v) and not one["neighbours"]
