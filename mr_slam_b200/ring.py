"""RING descriptor front end and matcher on the GPU (mirror of LoopDetection/src/RING_ros/util.py:174-200 `generate_RING`
and :362-374 `fast_corr`), batched: many scans per call instead of one python call per scan.

Data stays on the device between the stages (BEV -> Radon -> normalise -> FFT); descriptors are complex64 torch tensors
[B, 1, n_angles, det] like the reference's TIRING."""
import numpy as np
import torch

from ._lib import check, lib, ptr

NUM_RING = 120     # RING_ros/config.py:8-12
NUM_SECTOR = 120
NUM_HEIGHT = 1
MAX_LENGTH = 1
MAX_HEIGHT = 1


def _sync():
    check(lib().mrslam_device_synchronize())


def bev_batch(points_list, device="cuda:0", num_x=NUM_RING, num_y=NUM_SECTOR):
    """points_list: list of (N_i,3) float32 arrays already filtered+scaled by load_pc_infer.  Returns the max-height BEV
    images [B, num_x, num_y] (what generate_RING keeps of voxelocc's output: channel 2)."""
    offs = np.zeros(len(points_list) + 1, np.int64)
    for i, p in enumerate(points_list):
        offs[i + 1] = offs[i] + len(p)
    host = np.empty(3 * int(offs[-1]), np.float32)
    for i, p in enumerate(points_list):
        n = len(p)
        host[3 * offs[i]:3 * offs[i] + 3 * n] = np.asarray(p, np.float32)[:, :3].T.reshape(-1)  # pc.transpose().flatten()
    d_pts = torch.from_numpy(host).pin_memory().to(device, non_blocking=True)
    return bev_batch_device(d_pts, offs, num_x, num_y)


def bev_batch_device(d_pts_soa, offsets, num_x=NUM_RING, num_y=NUM_SECTOR):
    B = len(offsets) - 1
    bev = torch.empty((B, num_x, num_y), dtype=torch.float32, device=d_pts_soa.device)
    offs = np.ascontiguousarray(offsets, np.int64)
    torch.cuda.current_stream(d_pts_soa.device).synchronize()
    check(lib().mrslam_voxelocc_bev_device(ptr(d_pts_soa), offs.ctypes.data, B, MAX_LENGTH, MAX_HEIGHT, num_x, num_y, ptr(bev)))
    return bev


def scans_to_bev_device(d_scans, n_points=None, offsets=None, xy_range=70.0, z_range=30.0, num_x=NUM_RING, num_y=NUM_SECTOR):
    """RAW scans -> max-height BEV images in one kernel: load_pc_infer (util.py:91-112) + voxelocc channel 2 (util.py:174-188).
    d_scans: CUDA float32 [B, N, C>=3] (offsets None: every scan has N points) or [sum N_i, C] with `offsets` (B+1, points)."""
    if offsets is None:
        B, N, Cc = d_scans.shape
        offsets = np.arange(B + 1, dtype=np.int64) * N
    else:
        offsets = np.ascontiguousarray(offsets, np.int64)
        B, Cc = len(offsets) - 1, d_scans.shape[-1]
    bev = torch.empty((B, num_x, num_y), dtype=torch.float32, device=d_scans.device)
    torch.cuda.current_stream(d_scans.device).synchronize()
    check(lib().mrslam_ring_scans_to_bev_device(ptr(d_scans), offsets.ctypes.data, B, int(Cc), float(xy_range), float(z_range), MAX_LENGTH,
                                                MAX_HEIGHT, num_x, num_y, ptr(bev)))
    return bev


def ring_angles(device, n=NUM_RING):
    return torch.from_numpy(np.linspace(0, 2 * np.pi, n).astype(np.float32)).to(device)  # util.py:192 (endpoint inclusive)


def generate_RING_batch(bev, angles=None, interp="texture", fast_trig=True, want_sino=False):
    """bev: CUDA float32 [B, S, S].  Returns (sinogram [B,1,A,S] or None, TIRING complex64 [B,1,A,S])."""
    B, S, _ = bev.shape
    ang = angles if angles is not None else ring_angles(bev.device, S)
    A = ang.numel()
    tir = torch.empty((B, 1, A, S, 2), dtype=torch.float32, device=bev.device)
    sino = torch.empty((B, 1, A, S), dtype=torch.float32, device=bev.device) if want_sino else None
    torch.cuda.current_stream(bev.device).synchronize()
    check(lib().mrslam_ring_generate(ptr(bev.contiguous()), B, S, ptr(ang.contiguous()), A, 1 if interp == "texture" else 0,
                                     1 if fast_trig else 0, ptr(sino) if want_sino else None, ptr(tir)))
    _sync()
    return sino, torch.view_as_complex(tir)


def tiring(sino):
    """normalise + fft2(dim=-2, ortho) of sinograms [B, C, A, D] (float32 CUDA) -> complex64 [B, C, A, D]"""
    B, C, A, D = sino.shape
    tir = torch.empty((B, C, A, D, 2), dtype=torch.float32, device=sino.device)
    torch.cuda.current_stream(sino.device).synchronize()
    check(lib().mrslam_ring_tiring(ptr(sino.contiguous()), B, C, A, D, ptr(tir)))
    _sync()
    return torch.view_as_complex(tir)


def fast_corr_batch(query, database, want_corr=False):
    """query: complex64 [C, A, D]; database: complex64 [N, C, A, D] (both CUDA).  Returns (dist[N], angle[N]) like fast_corr."""
    N, C, A, D = database.shape
    q = torch.view_as_real(query.contiguous()).contiguous()
    db = torch.view_as_real(database.contiguous()).contiguous()
    dist = torch.empty(N, dtype=torch.float32, device=db.device)
    ang = torch.empty(N, dtype=torch.int32, device=db.device)
    corr = torch.empty((N, A), dtype=torch.float32, device=db.device) if want_corr else None
    torch.cuda.current_stream(db.device).synchronize()
    check(lib().mrslam_ring_fast_corr(ptr(q), ptr(db), N, C, A, D, ptr(dist), ptr(ang), ptr(corr) if want_corr else None))
    _sync()
    return (dist, ang, corr) if want_corr else (dist, ang)


def calculate_row_shift(shift, num_ring=120):
    """util.py:377-383: circular row shift that compensates the matched yaw (in grid units)"""
    return -shift if shift < num_ring // 2 else shift - num_ring


def solve_translation(query, positive, rot_angle, device=None, row_shift=0, want_shifts=False):
    """solve_translation(query, positive, rot_angle, device) of RING/util.py:388-423: query / positive are CUDA float32 sinograms
    [B, H, W] (B = channels).  row_shift folds main_RING.py's torch.roll(RING_matched, row_shift, dims=1) into the call (pass the
    unrolled RING_matched).  Returns (x, y, error) as numpy scalars like the reference (and the per-row shifts b[H] on request)."""
    q, p = query.contiguous().float(), positive.contiguous().float()
    B, H, W = q.shape
    out = np.zeros(3, np.float32)
    b = torch.empty(H, dtype=torch.float32, device=q.device) if want_shifts else None
    torch.cuda.current_stream(q.device).synchronize()
    check(lib().mrslam_ring_solve_translation(ptr(q), ptr(p), B, H, W, float(rot_angle), int(row_shift), ptr(b) if want_shifts else None,
                                              out.ctypes.data))
    x, y, err = np.float32(out[0]), np.float32(out[1]), np.float32(out[2])
    return (x, y, err, b) if want_shifts else (x, y, err)


def build_neighbors_NN(pointcloud, k):
    """build_neighbors_NN(pointcloud, k) of util.py:163-170 on the GPU (exact tiled kNN + covariance / eigenvalue kernel).
    Returns (indices [n,k] int32, covs [n,3,3], entropy [n], eigens_ [n,5], vectors_ = []) like the reference."""
    pc = np.ascontiguousarray(np.asarray(pointcloud)[:, :3], np.float32)
    n = pc.shape[0]
    idx = np.zeros((n, k), np.int32)
    cov = np.zeros((n, 3, 3), np.float32)
    ent = np.zeros(n, np.float32)
    eig = np.zeros((n, 5), np.float32)
    check(lib().mrslam_ringpp_build_neighbors(ptr(pc), n, 3, int(k), ptr(idx), ptr(cov), ptr(ent), ptr(eig)))
    return idx, cov, ent, eig, []


def calc_rel_ori(new_disco, old_discos, want_loc=False):
    """GlobalManager::calcRelOri (global_manager.cpp:2719-2762), batched over the old signatures: new_disco = (real, imag) CUDA float32
    [H, W], old_discos = (real, imag) CUDA float32 [N, H, W].  Returns rel_angle [N] in degrees (and the flat arg-max)."""
    ra, ia = (t.contiguous().float() for t in new_disco)
    rb, ib = (t.contiguous().float() for t in old_discos)
    N, H, W = rb.shape
    ang = torch.empty(N, dtype=torch.float32, device=rb.device)
    loc = torch.empty(N, dtype=torch.int32, device=rb.device) if want_loc else None
    torch.cuda.current_stream(rb.device).synchronize()
    check(lib().mrslam_disco_rel_ori(ptr(ra), ptr(ia), ptr(rb), ptr(ib), N, H, W, ptr(ang), ptr(loc) if want_loc else None))
    _sync()
    return (ang, loc) if want_loc else ang

