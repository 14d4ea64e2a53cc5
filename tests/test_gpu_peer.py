"""The fused detection gather (viddet_b200.dist.PeerGather + output mirrors) across TWO PROCESSES.  CUDA IPC needs distinct
processes, not distinct GPUs: both ranks use cuda:0 here (the multi-GPU bench maps the same buffers over NVLink), gloo carries
the handles."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        import viddet_b200
        from viddet_b200 import dist as vdist
        from tests.util import make_pred_weights, make_tips
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
        vdist.init_from_env(backend="gloo")
        torch.cuda.set_device(0)
        C, B, size = 20, 2, 320
        rng = np.random.RandomState(5)
        ws, bs = make_pred_weights(rng, C, bias_scale=0.1)
        head = viddet_b200.YOLOV3Head(C)
        for o, w, b in zip(head.yolo_outputs, ws, bs):
            o.prediction.set_data(torch.from_numpy(w), torch.from_numpy(b))
        head.set_nms(0.45, 400, 100)
        n = B * 100
        pg = vdist.PeerGather(6 * n)
        own = pg.slot
        out = (own[:n].view(B, 100, 1), own[n:2 * n].view(B, 100, 1), own[2 * n:6 * n].view(B, 100, 4))
        tips = [torch.from_numpy(t).cuda() for t in make_tips(np.random.RandomState(100 + rank), B, size=size)]   # rank-specific frames
        sess = head.session(tips, out=out, mirrors=pg.deltas)
        sess.run(); sess.run()
        torch.cuda.synchronize()
        dist.barrier()
        # what every rank SHOULD hold: each rank's own result, recomputed locally for all ranks
        ok = True
        for r in range(world):
            t_r = [torch.from_numpy(t).cuda() for t in make_tips(np.random.RandomState(100 + r), B, size=size)]
            ids, scores, boxes = head(t_r)
            exp = torch.cat([ids.reshape(-1), scores.reshape(-1), boxes.reshape(-1)])
            ok = ok and torch.equal(pg.gathered[r][:6 * n].view(torch.int32), exp.view(torch.int32))
        torch.cuda.synchronize()
        dist.barrier()
        pg.close()
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:                                    # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()))


def test_peer_gather_two_processes():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, ok, msg in res:
        assert ok, "rank %d: %s" % (rank, msg)
