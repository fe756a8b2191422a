"""Multi-GPU parity check of the partitioned path (run under torchrun on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py [nx|mesh]

Every rank builds a partitioned context of the same mesh; rank 0 also steps an ordinary single-GPU context of the whole
mesh.  tests/dist_parity.py ASSERTS: iteration counts within +-3 over a trajectory from rest and the state <= 1e-4 at the
shipped eps = 1e-6; from equal states every rank's owned rows of rhs and Keff bit-exact; after a tight solve
(eps = 1e-12) displacement and velocity <= 1e-8.  Then a second trajectory WITHOUT any host-side synchronisation between the
steps (the production call pattern: back-to-back fb_step) must reproduce the first one bit for bit — the cross-solve reuse
of the peer-memory slots (double-buffered by solve parity, fb_pcg_common.cuh) is what that exercises."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import fembrain_b200 as fb  # noqa: E402
from tests import cases, dist_parity  # noqa: E402


def main():
    arg = sys.argv[1] if len(sys.argv) > 1 else "12"
    nx = int(arg) if arg.isdigit() else arg
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(fb.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    comm_id = bytes(idt.cpu().numpy().tobytes())

    if isinstance(nx, int):
        v, t, fixed, load = cases.cube_case(nx)
    else:  # a mesh from the reference's data directory (unstructured numbering: the partition is cut from a Cuthill-McKee ordering)
        v, t, fixed = cases.golden_mesh(nx)
        load = int(np.argmax(v[:, 1] * 1000 + v[:, 0]))
    r = 3 * len(v)
    f = cases.point_load(r, load)
    part = fb.Simulation(v, t, fixed, partition=(rank, world, comm_id), device=local)
    assert part.r == r
    b, e = part.partition_range()
    part.set_external_forces(f)
    ref = fb.Simulation(v, t, fixed, device=local) if rank == 0 else None
    if ref is not None:
        ref.set_external_forces(f)
    res = dist_parity.partition_parity(part, ref, rank, world, r, traj_steps=3, log=lambda *a: print(*a, flush=True))
    ok = res["ok"]
    # back-to-back steps, no collective or host sync of ours between them: same trajectory, bit for bit, twice
    finals = []
    for rep in range(2):
        part.reset_to_rest()
        its = []
        for _ in range(6):
            part.do_timestep()
            its.append(int(part.last_cg_iterations))
        q, qv = part.get_state_owned()
        finals.append((its, q.copy(), qv.copy()))
    same = finals[0][0] == finals[1][0] and np.array_equal(finals[0][1], finals[1][1]) and np.array_equal(finals[0][2], finals[1][2])
    flag = torch.tensor([1 if (ok and same) else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    good = int(flag[0]) == 1
    if rank == 0:
        res["back_to_back_reproducible"] = bool(same)
        print("DIST_PARITY " + json.dumps(res), flush=True)
        print("DIST_CHECK", "OK" if good else "FAILED", f"world={world} mesh={nx} rows [{b},{e}) peer_memory={part.peer_memory} reordered={part.reordered}", flush=True)
    part.close()
    dist.destroy_process_group()
    sys.exit(0 if good else 1)


if __name__ == "__main__":
    main()
