"""Multi-GPU check of the halo exchange over NCCL (run under torchrun): every rank clusters its tile with
tiles.tile_dbscan (NCCL send/recv of the halo, all-gathers, all-reduce of the cluster table); rank 0 compares the
concatenated labels with ONE device run on the concatenated candidates and with scikit-learn.
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/nccl_halo_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from pointcloudhookup_b200 import device as dv, tiles as tl
import halo_oracle as ho

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
EPS, MINPTS = 8.0, 20
tiles = ho.corridor_candidates(41, world, per_tile=40000)
comm = tl.TorchComm(dev)
res = tl.tile_dbscan(torch.from_numpy(tiles[rank]).to(dev), (1.0, 0.0), EPS, MINPTS, comm)
lab = res.labels.cpu().numpy()
parts = comm.all_gather_np(lab)
if rank == 0:
    from sklearn.cluster import DBSCAN
    allp = np.concatenate(tiles)
    got = np.concatenate(parts)
    one = dv.dbscan_chunked(torch.from_numpy(allp).to(dev), EPS, MINPTS, chunk=len(allp))
    ref = DBSCAN(eps=EPS, min_samples=MINPTS, algorithm="ball_tree").fit(allp).labels_.astype(np.int32)
    ok = np.array_equal(got, one.labels.cpu().numpy()) and np.array_equal(got, ref) and res.n_clusters == one.n_clusters
    ok = ok and np.array_equal(res.stats["count"], one.stats["count"]) and np.array_equal(res.stats["min"], one.stats["min"])
    print(f"nccl halo check world={world}: K={res.n_clusters} halo sent={res.sent} recv={res.halo} p2p_bytes={comm.bytes_p2p} "
          f"gather_bytes={comm.bytes_gather} -> {'OK' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        sys.exit(1)
dist.barrier()
dist.destroy_process_group()
