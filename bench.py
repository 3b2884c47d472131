#!/usr/bin/env python
"""bench.py — chunks/sec of the MVPNet forward (8192 pts, 5 views 160x120) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of synthetic chunks per GPU (BASELINE config 4's
per-rank shard: 32 chunks x 8192 points x 5 views of 160x120): depth unprojection + 2D->3D 3-NN,
UNet-ResNet34 on the views (this package's tcgen05 convolutions), FeatureAggregation, PN2SSG, and (N > 1) one NCCL
all-gather of the logits.

  value  chunks/s with the step's inputs already resident in HBM (CUDA events, max over ranks)
  e2e    chunks/s through the public API (engine.PipelinedForward.submit) from pinned HOST buffers: H2D of every
         step's inputs and the D2H read of its logits are inside the timed region (copy streams, double-buffered)
  roofline      the dominant kernel of this package inside the step, timed live with CUDA events
  cpu_baseline  the same forward on the host CPU cores (PyTorch CPU modules of the same architecture +
                the C oracle for the six extension ops + numpy/scikit-learn for unprojection / k-NN, as the
                reference's DataLoader workers do), on a bounded sample

--impl reference runs only that CPU arm (the reference has no CPU implementation of its CUDA ops and
its sources cannot travel to the GPU box; kind = "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_POINTS, NUM_VIEWS, H, W, NUM_CLASSES, KNN = 8192, 5, 120, 160, 20, 3
METRIC = 'chunks/sec MVPNet fwd (8192 pts, 5 views 160x120)'


NUMA_NOTE = None


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        a, _, b = part.strip().partition('-')
        if a.isdigit():
            cpus.update(range(int(a), int(b if b.isdigit() else a) + 1))
    return cpus


def bind_to_gpu_numa(local_rank):
    """Best effort: pin this process to the CPUs next to its GPU BEFORE any pinned host buffer is allocated (first touch
    places the pages there).  With 8 ranks on one box every rank otherwise pins ~40 MB per step on whatever node the
    launcher started it on (e2e scaling 0.83 at N = 8 with device scaling 0.96).  Sources, in order: the GPU's
    sysfs numa_node, then the 'CPU Affinity' column of `nvidia-smi topo -m` (virtualised boxes report numa_node = -1).
    What was found is kept in NUMA_NOTE and printed in the JSON line."""
    global NUMA_NOTE
    import re
    allowed = os.sched_getaffinity(0)
    try:
        bus = subprocess.run(['nvidia-smi', '-i', str(local_rank), '--query-gpu=pci.bus_id', '--format=csv,noheader'],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if len(bus.split(':')[0]) == 8:
            bus = bus[4:]
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read())
        if node >= 0:
            cpus = _cpulist(open('/sys/devices/system/node/node%d/cpulist' % node).read()) & allowed
            if cpus:
                os.sched_setaffinity(0, cpus)
                NUMA_NOTE = 'sysfs numa_node %d, %d of %d allowed cpus' % (node, len(cpus), len(allowed))
                return node
            NUMA_NOTE = 'sysfs numa_node %d has no cpu this process may use (%d allowed)' % (node, len(allowed))
            return None
        note = 'sysfs numa_node = %d for %s' % (node, bus)
    except Exception as e:
        note = 'sysfs lookup failed: %s' % type(e).__name__
    try:
        topo = subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True, timeout=20).stdout
        lines = [re.sub(r'\x1b\[[0-9;]*m', '', ln) for ln in topo.splitlines()]
        head = next(ln for ln in lines if 'CPU Affinity' in ln).split('\t')
        col = [h.strip() for h in head].index('CPU Affinity')
        row = next(ln for ln in lines if ln.split('\t')[0].strip() == 'GPU%d' % local_rank).split('\t')
        cpus = _cpulist(row[col]) & allowed
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            NUMA_NOTE = note + '; nvidia-smi topo CPU affinity %s -> %d of %d allowed cpus' % (row[col].strip(), len(cpus), len(allowed))
            numa_col = [h.strip() for h in head].index('NUMA Affinity') if 'NUMA Affinity' in [h.strip() for h in head] else -1
            return int(row[numa_col].strip()) if numa_col >= 0 and row[numa_col].strip().isdigit() else -1
        NUMA_NOTE = note + '; nvidia-smi topo CPU affinity "%s" does not narrow the %d allowed cpus' % (row[col].strip(), len(allowed))
    except Exception as e:
        NUMA_NOTE = note + '; nvidia-smi topo lookup failed: %s' % type(e).__name__
    return None


# ------------------------------------------------------------------------------------------------
def build_model(device):
    from mvpnet_b200 import synthetic
    from mvpnet_b200.modules import MVPNet3D, PN2SSG
    from mvpnet_b200.unet import UNetResNet34
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        net2d = UNetResNet34(NUM_CLASSES, p=0.5, pretrained=False)
    model = MVPNet3D(net2d, None, PN2SSG(64, NUM_CLASSES), in_channels=64, mlp_channels=(64, 64, 64),
                     reduction='sum', use_relation=True)
    synthetic.fill_parameters(model, seed=6)
    return model.eval().to(device)


def make_host_batch(seeds, pin):
    from mvpnet_b200 import synthetic
    from mvpnet_b200.data import invert_intrinsics
    chunks = [synthetic.make_chunk(seed=s, num_points=NUM_POINTS, num_views=NUM_VIEWS, h=H, w=W) for s in seeds]
    # Inputs travel in the formats the dataset stores (scannet_2d3d.py:229-251): colour as uint8 HWC, depth as uint16
    # millimetres (held in an int16 tensor: torch has no first-class uint16); the device decodes them (engine.decode_stored_inputs).
    rgb = np.stack([np.clip(c['images'].transpose(0, 2, 3, 1) * 40 + 128, 0, 255).astype(np.uint8) for c in chunks])
    host = {
        'images_u8': torch.from_numpy(rgb),                                               # (b, nv, h, w, 3) uint8
        'depth_mm': torch.from_numpy(np.stack([c['depth_mm'] for c in chunks]).view(np.int16)),   # (b, nv, h, w) uint16 bits
        'pose': torch.from_numpy(np.stack([c['pose'] for c in chunks])),
        'cam_inv': torch.from_numpy(np.stack([np.broadcast_to(invert_intrinsics(c['cam_matrix']), (NUM_VIEWS, 3, 3)) for c in chunks]).copy()),
        'points': torch.from_numpy(np.stack([c['points'] for c in chunks])),          # (b, np, 3)
        'chunk_box': torch.from_numpy(np.stack([c['chunk_box'] for c in chunks])),
    }
    if pin:
        host = {k: v.pin_memory() for k, v in host.items()}
    return host, chunks


def hot_path(model, dev, overlap=True):
    """One step on device-resident inputs (raw depth / pose / images / points): returns seg_logit (b, 20, np)."""
    batch = {'images_u8': dev['images_u8'], 'points': dev['points_cm'], 'depth_mm': dev['depth_mm'], 'pose': dev['pose'],
             'cam_inv': dev['cam_inv'], 'chunk_box': dev['chunk_box'], 'k': KNN}
    return model.fast_forward(batch, overlap=overlap)['seg_logit']


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': float(self.rows[0][1]) if self.rows[0][1].replace('.', '').isdigit() else None,
                'samples': len(self.rows), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# algorithmic bytes / work per launch of each kernel, per chunk
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes_per_chunk():
    P = NUM_VIEWS * H * W
    n = [8192, 2048, 512, 128, 32]
    c_sa_in = [64, 64, 128, 256]
    c_sa_out = [64, 128, 256, 512]
    out = {
        'unproject': P * 4 + P * (12 + 24 + 1),
        'knn_pixels': P * 25 + NUM_POINTS * 24 + NUM_POINTS * KNN * 16,
        'feature_aggregation': NUM_POINTS * KNN * (64 * 4 + 12 + 8) + NUM_POINTS * 12 + NUM_POINTS * 64 * 4,
    }
    for i in range(4):
        N, M = n[i], n[i + 1]
        out['fps%d' % (i + 1)] = N * 12 + M * 8
        out['ball_query%d' % (i + 1)] = M * 12 + N * 12 + M * 32 * 8
        out['set_abstraction%d' % (i + 1)] = N * c_sa_in[i] * 4 + N * 12 + M * 32 * 8 + M * 12 + M * c_sa_out[i] * 4
    fp_cs, fp_cd, fp_out = [512, 256, 256, 128], [256, 128, 64, 0], [256, 256, 128, NUM_CLASSES]
    for i in range(4):
        Ns, Nd = n[4 - i], n[3 - i]
        out['knn_distance%d' % (i + 1)] = Nd * 12 + Ns * 12 + Nd * 3 * 12
        out['feature_propagation%d' % (i + 1)] = Ns * fp_cs[i] * 4 + Nd * 36 + Nd * fp_cd[i] * 4 + Nd * fp_out[i] * 4
    return out


# multiply-adds of the UNet-ResNet34 per 128x160 (padded) view, by kernel family (counted layer by layer from
# unet_resnet34.py:9-125): 3x3/stride-1 convolutions 8114 M, the other layers (7x7 stem, three stride-2 3x3, three 1x1
# down-samples, four 2x2 transposed) 715 M
NET2D_MMAC_PER_VIEW = {'net_2d/conv3x3': 8114.0, 'net_2d/conv_general': 715.0}


def algorithmic_gflop_per_chunk():
    """Contraction flops per chunk: MLP chains and the 2D network's convolutions."""
    conv = {k: 2 * v * NUM_VIEWS / 1e3 for k, v in NET2D_MMAC_PER_VIEW.items()}
    conv['net_2d'] = sum(conv.values())
    return {**conv, 'feature_aggregation': 0.617, 'set_abstraction1': 0.684, 'set_abstraction2': 0.543, 'set_abstraction3': 0.540,
            'set_abstraction4': 0.538, 'feature_propagation1': 0.067, 'feature_propagation2': 0.168,
            'feature_propagation3': 0.470, 'feature_propagation4': 0.805 + 0.310}


def stage_bound(name):
    if name == 'net_2d':
        return 'tensor (tcgen05 convolutions, bf16 hi/lo x3; sum of the net_2d/* stages)'
    if name in ('net_2d/conv3x3', 'net_2d/conv_general'):
        return 'tensor (tcgen05, bf16 hi/lo x3)'
    if name == 'net_2d/pool_unfold':
        return 'hbm'
    if name in ('unproject',):
        return 'hbm'
    if name == 'feature_aggregation':
        return 'hbm (pixel-feature gather) + tensor'
    if name.startswith('set_abstraction') or name.startswith('feature_propagation'):
        return 'tensor (tcgen05, bf16 hi/lo x3)'
    if name.startswith('fps'):
        return 'serial latency / issue (M-1 dependent arg-max steps, one CTA per cloud)'
    if name == 'knn_pixels':
        return 'fp64 CUDA-core compute + latency (exact grid search)'
    return 'fp32 CUDA-core compute (exhaustive search)'


# kernels of this package per step (counted in profiles/r1_launches_*: unproject 1, pixel k-NN grid 5, FPS 4,
# ball query 4 + grid 5, 3-NN 4 + grid 5, fused FA/SA/FP 9, 2D network: 3x3 convolutions 33, other layers 11,
# stem unfold 1, max-pool 1); ATen adds 11 small copies / gathers
KERNELS_PER_STEP = 83
# dram__bytes_read.sum + dram__bytes_write.sum per launch (MB) from the committed ncu capture (profiles/r1_ncu_*.md)
NCU_TRAFFIC_MB = {'net_2d/conv3x3': 285.7}     # profiles/r1_ncu_full_final.md: 9429 MB over the 33 launches of one step


def tensor_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path)).get('bf16_tflops', 1590.0))
    return 1590.0


def tensor_peak_sustained():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p.get('bf16_tflops_sustained', p.get('bf16_tflops', 1400.0)))
    return 1400.0


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return p['hbm_gbs'], 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------------
# CPU arm: reference architecture on the host cores
# ------------------------------------------------------------------------------------------------
class _OracleExt:
    """Stands in for the CUDA extension so that mvpnet_b200.modules runs on CPU tensors — used ONLY by the
    cpu_baseline / --impl reference legs of this benchmark."""

    def __init__(self):
        import oracle
        for k, v in oracle.ext_modules().items():
            setattr(self, k, v)


def cpu_forward_factory():
    import oracle
    from sklearn.neighbors import NearestNeighbors
    import mvpnet_b200.ops._util as util
    inst = _OracleExt()
    util.ext = lambda: inst
    for mod in ('fps', 'ball_query', 'group_points', 'knn_distance', 'interpolate'):
        m = __import__('mvpnet_b200.ops.' + mod, fromlist=['ext'])
        m.ext = util.ext
    model = build_model('cpu')

    def run(chunk):
        # data side exactly as scannet_2d3d.py:255-313: numpy unprojection (oracle restatement) + sklearn ball tree
        ci = np.linalg.inv(chunk['cam_matrix'][:3, :3])
        xyz64, xyz32, mask = oracle.unproject(chunk['depth'], ci, chunk['pose'], chunk['chunk_box'])
        valid = np.nonzero(mask.reshape(-1))[0]
        nbrs = NearestNeighbors(n_neighbors=KNN, algorithm='ball_tree').fit(xyz64.reshape(-1, 3)[valid])
        _, knn = nbrs.kneighbors(chunk['points'])
        knn = valid[knn]
        # colour decoded from the stored uint8 exactly as the GPU arm's device kernel does (scannet_2d3d.py:229-246)
        from mvpnet_b200.engine import IMAGE_MEAN, IMAGE_STD
        u8 = np.clip(chunk['images'].transpose(0, 2, 3, 1) * 40 + 128, 0, 255).astype(np.uint8)
        img = ((u8.astype(np.float32) / np.float32(255.0) - np.asarray(IMAGE_MEAN, np.float32)) / np.asarray(IMAGE_STD, np.float32)).transpose(0, 3, 1, 2)
        batch = {'images': torch.from_numpy(np.ascontiguousarray(img))[None], 'image_xyz': torch.from_numpy(xyz32)[None],
                 'knn_indices': torch.from_numpy(knn.astype(np.int64))[None],
                 'points': torch.from_numpy(np.ascontiguousarray(chunk['points'].T))[None]}
        with torch.no_grad():
            return model(batch)['seg_logit']
    return run


def cpu_arm(steps, warmup, chunks_per_step=1):
    from mvpnet_b200 import synthetic
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    run = cpu_forward_factory()
    chunks = [synthetic.make_chunk(seed=1000 + i) for i in range(chunks_per_step)]
    for _ in range(warmup):
        run(chunks[0])
    t0 = time.perf_counter()
    for _ in range(steps):
        for c in chunks:
            run(c)
    dt = time.perf_counter() - t0
    n = steps * chunks_per_step
    return {'value': n / dt, 'unit': 'chunks/s', 'cores': cores, 'kind': 'port',
            'sample': '%d chunk forward(s) of the same workload (8192 pts, 5 views) on %d host threads: PyTorch CPU modules + C oracle ops + sklearn ball-tree k-NN' % (n, cores),
            'ms_per_chunk': dt / n * 1e3}


def reference_kernel_times(points_pm, iters=3):
    """Device time of the REFERENCE's own CUDA kernels (oracle/_ref, built unmodified for sm_100a by
    oracle/build_ref.py) at the model's level-1 / level-4 shapes, next to this package's kernels on the same
    inputs.  Reported for context ("the reference algorithm on B200"); skipped when oracle/_ref is absent."""
    import glob
    import importlib.util
    import mvpnet_b200
    ext = mvpnet_b200.load_ext()
    ref = {}
    for name in ('fps_cuda', 'ball_query_cuda', 'group_points_cuda', 'knn_distance_cuda', 'interpolate_cuda'):
        hits = glob.glob(os.path.join(ROOT, 'oracle', '_ref', name + '*.so'))
        if not hits:
            return None
        spec = importlib.util.spec_from_file_location(name, hits[0])
        ref[name] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref[name])

    def t(fn):
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters

    b = points_pm.size(0)
    idx = ext.fps_cuda.farthest_point_sample(points_pm, 2048)
    cent = torch.gather(points_pm, 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    nbr = ext.ball_query_cuda.ball_query(cent, points_pm, 0.1, 32)
    feat = torch.randn(b, 64, 8192, device=points_pm.device)
    ki, kd = ext.knn_distance_cuda.knn_distance(points_pm, cent, 3)
    w = torch.rand(b, 8192, 3, device=points_pm.device)
    f2 = torch.randn(b, 128, 2048, device=points_pm.device)
    out = {}
    for label, mine, theirs in [
            ('fps 8192->2048', lambda: ext.fps_cuda.farthest_point_sample(points_pm, 2048), lambda: ref['fps_cuda'].farthest_point_sample(points_pm, 2048)),
            ('ball_query 2048x8192 r0.1 K32', lambda: ext.ball_query_cuda.ball_query(cent, points_pm, 0.1, 32), lambda: ref['ball_query_cuda'].ball_query(cent, points_pm, 0.1, 32)),
            ('group_points 64ch 2048x32', lambda: ext.group_points_cuda.group_points_forward(feat, nbr), lambda: ref['group_points_cuda'].group_points_forward(feat, nbr)),
            ('knn_distance 8192x2048', lambda: ext.knn_distance_cuda.knn_distance(points_pm, cent, 3), lambda: ref['knn_distance_cuda'].knn_distance(points_pm, cent, 3)),
            ('interpolate 128ch 2048->8192', lambda: ext.interpolate_cuda.interpolate_forward(f2, ki, w), lambda: ref['interpolate_cuda'].interpolate_forward(f2, ki, w))]:
        out[label] = {'b200_ms': round(t(mine), 4), 'reference_kernel_ms': round(t(theirs), 4)}
        out[label]['speedup'] = round(out[label]['reference_kernel_ms'] / out[label]['b200_ms'], 2)
    return out


def _time_ms(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def count_kernel_launches(fn):
    """Kernel launches of one eager call of `fn`, from the CUPTI activity records (torch.profiler): (launches of this
    package's kernels, all launches).  Counted, not assumed."""
    from torch.profiler import ProfilerActivity, profile
    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    mine = total = 0
    for ev in prof.events():
        if str(ev.device_type).endswith('CUDA') and not ev.name.startswith('Memcpy') and not ev.name.startswith('Memset'):
            total += 1
            if 'mvp::' in ev.name or 'tcc::' in ev.name or 'tc2::' in ev.name or 'tc::' in ev.name:
                mine += 1
    return mine, total


def ncu_traffic_mb():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (MB) of the dominant kernel families, read from the
    committed ncu summary (profiles/r2_ncu_traffic.json, written by tools/ncu_traffic.py from an `ncu --set full` capture
    of this same command); None when the file is absent."""
    path = os.path.join(ROOT, 'profiles', 'r2_ncu_traffic.json')
    if not os.path.exists(path):
        return {}, None
    return json.load(open(path)), 'profiles/r2_ncu_traffic.json'


def bench_extras(model, device, dev, cpg):
    """Sub-lines the headline does not cover (VERDICT r1 'measurement completeness'): batch-1 latency, BASELINE config 2
    (PN2SSG forward, and the training step forward + backward at batch 1 / 32 with its loss checked against the reference
    golden), BASELINE config 5 (whole-scene PN2SSG on ~200 k points and the reference's 3 x 32768 test shape: time, peak
    memory, stages), the chunked whole-scene pipeline in scenes/s, and an in-bench parity check of the timed batch."""
    import mvpnet_b200
    from mvpnet_b200 import engine, scene, synthetic, train
    from mvpnet_b200.modules import PN2SSG
    out = {}
    with torch.no_grad():
        # ---- parity of the timed path: fused forward of two chunks of the timed batch == op-by-op module composition
        sub = {k: (v[:2] if torch.is_tensor(v) and v.dim() > 0 and v.size(0) == cpg else v) for k, v in dev.items()}
        fast = hot_path(model, sub)
        dec = engine.decode_stored_inputs({'images_u8': sub['images_u8'], 'depth_mm': sub['depth_mm']})
        rg = engine._data_side({'depth': dec['depth'], 'pose': sub['pose'], 'cam_inv': sub['cam_inv'], 'chunk_box': sub['chunk_box']},
                               sub['points'], KNN)
        slow = model({'images': dec['images'], 'image_xyz': rg['image_xyz'], 'knn_indices': rg['knn_indices'], 'points': sub['points_cm']})['seg_logit']
        # the decode kernels against the host conversions they replace (float32, same operation order): bit-exact
        u8 = sub['images_u8'].cpu().numpy().astype(np.float32) / np.float32(255.0)
        want = (u8 - np.asarray(engine.IMAGE_MEAN, np.float32)) / np.asarray(engine.IMAGE_STD, np.float32)
        if not np.array_equal(dec['images'].cpu().numpy(), want.transpose(0, 1, 4, 2, 3)):
            raise RuntimeError('bench: device colour decode differs from the host conversion')
        if not np.array_equal(dec['depth'].cpu().numpy(), sub['depth_mm'].cpu().numpy().view(np.uint16).astype(np.float32) / np.float32(1000.0)):
            raise RuntimeError('bench: device depth decode differs from the host conversion')
        err = float((fast - slow).abs().max() / slow.abs().max())
        out['parity_timed_batch'] = {'fused_vs_op_by_op_rel_err': err, 'tolerance': 1e-4, 'ok': err < 1e-4}
        if err >= 1e-4:
            raise RuntimeError('bench: fused forward of the timed batch differs from the op-by-op path (%.3g)' % err)
        # ---- batch-1 latency (BASELINE config 3 is literally one chunk; the reference's test loop is batch 1)
        one = {k: (v[:1].contiguous() if torch.is_tensor(v) and v.dim() > 0 and v.size(0) == cpg else v) for k, v in dev.items()}
        b1 = {'images_u8': one['images_u8'], 'points': one['points_cm'], 'depth_mm': one['depth_mm'], 'pose': one['pose'], 'cam_inv': one['cam_inv'],
              'chunk_box': one['chunk_box'], 'k': KNN}
        g1 = engine.GraphedForward(model, b1)
        ms = _time_ms(lambda: g1(b1), 30, warm=5)
        out['b1_latency'] = {'ms_per_chunk': round(ms, 4), 'chunks_per_s': round(1e3 / ms, 1), 'mode': 'one CUDA graph replay per chunk, inputs resident'}
        # ---- config 2: PN2SSG forward (eval, fused)
        c2 = {}
        net = synthetic.fill_parameters(PN2SSG(64, NUM_CLASSES, dropout_prob=0.0), seed=5).to(device)
        for b in (1, 32):
            pts, feat, label, weight = synthetic.train_batch(b)
            data = {'points': torch.from_numpy(pts.transpose(0, 2, 1).copy()).to(device), 'feature': feat.to(device)}
            net.eval()
            c2['forward_eval_fused_b%d_ms' % b] = round(_time_ms(lambda: net.fast_forward(data), 10), 4)
        out['config2_pn2ssg'] = c2
    # ---- config 2: training step (train-mode BatchNorm, SegLoss, metrics, deterministic backward), loss vs the reference golden
    for b in (1, 32):
        pts, feat, label, weight = synthetic.train_batch(b)
        data = {'points': torch.from_numpy(pts.transpose(0, 2, 1).copy()).to(device), 'feature': feat.to(device).requires_grad_(True),
                'seg_label': label.to(device)}
        net.train()
        loss_fn = train.SegLoss(weight=weight.to(device))

        def step():
            net.zero_grad(set_to_none=True)
            return train.train_step(net, loss_fn, data, metrics=(train.SegAccuracy(), train.SegIoU(NUM_CLASSES)))[1]['seg_loss']
        loss = float(step().item())
        gold = os.path.join(ROOT, 'tests', 'golden', 'pn2_train_b%d.npz' % b)
        ref_loss = float(np.load(gold)['loss']) if os.path.exists(gold) else None
        ms = _time_ms(step, 5, warm=1)
        out['config2_pn2ssg']['train_step_fwd_bwd_b%d' % b] = {
            'ms': round(ms, 3), 'chunks_per_s': round(b * 1e3 / ms, 1), 'loss': loss, 'reference_loss_fp64': ref_loss,
            'loss_rel_err': None if ref_loss is None else abs(loss - ref_loss) / ref_loss,
            'path': 'op-by-op modules on this package\'s kernels (batch-statistics BatchNorm), SegLoss + metrics kernels, deterministic scatter backward'}
        if ref_loss is not None and abs(loss - ref_loss) > 1e-4 * ref_loss:
            raise RuntimeError('bench: training-step loss differs from the reference golden')
    del net
    # ---- config 5: whole-scene PN2SSG
    with torch.no_grad():
        snet = synthetic.fill_parameters(PN2SSG(0, NUM_CLASSES, num_centroids=(8192, 2048, 512, 128)), seed=8).eval().to(device)
        rng = np.random.RandomState(0)
        c5 = {}
        for label, b, n in (('200k_points_b1', 1, 200000), ('reference_test_shape_b3_x_32768', 3, 32768)):
            p = rng.uniform([0, 0, 0], [6.0, 8.0, 2.7], (b, n, 3))
            sel = rng.rand(b, n)
            p[sel < 0.4, 2] = 0.0
            p[(sel >= 0.4) & (sel < 0.6), 0] = 0.0
            p[(sel >= 0.6) & (sel < 0.8), 1] = 8.0
            p = (p + rng.randn(b, n, 3) * 0.005).astype(np.float32)
            batch = {'points': torch.from_numpy(np.ascontiguousarray(p.transpose(0, 2, 1))).to(device)}
            snet.fast_forward(batch)
            torch.cuda.synchronize()
            torch.cuda.reset_peak_memory_stats()
            with engine.profile() as prof:
                ms = _time_ms(lambda: snet.fast_forward(batch), 2, warm=0)
            st = {k: round(float(np.median(v)), 3) for k, v in prof.summary().items()}
            c5[label] = {'forward_ms': round(ms, 3), 'points_per_s': round(b * n * 1e3 / ms), 'peak_mem_MB': round(torch.cuda.max_memory_allocated() / 2 ** 20, 1),
                         'stages_ms': st}
        out['config5_whole_scene_pn2ssg'] = c5
        # ---- chunked whole-scene pipeline: scene2chunks -> per-chunk forward (batch 1, variable size) -> votes -> labels
        tiles = [synthetic.room_points(5000, seed=900 + i)[0] + np.array([(i % 3) * 1.9, (i // 3) * 1.9, 0.0], np.float32) for i in range(12)]
        sp = torch.from_numpy(np.concatenate(tiles).astype(np.float32)).to(device)
        cnet = synthetic.fill_parameters(PN2SSG(0, NUM_CLASSES), seed=9).eval().to(device)

        def run_scene():
            idx = scene.scene2chunks_legacy(sp, chunk_size=(1.5, 1.5), stride=0.5, thresh=1000, margin=(0.2, 0.2))
            fwd = lambda batch: cnet.fast_forward(batch)['seg_logit']
            return len(idx), scene.chunked_scene_forward(fwd, idx, lambda ind: {'points': sp.index_select(0, ind).t().contiguous()[None]},
                                                         sp.size(0), NUM_CLASSES, device)
        nchunks, _ = run_scene()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            run_scene()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        out['scene_pipeline'] = {'scenes_per_s': round(1.0 / dt, 3), 'ms_per_scene': round(dt * 1e3, 1), 'chunks_per_scene': nchunks,
                                 'points_per_scene': int(sp.size(0)),
                                 'what': 'scene2chunks (1.5 m windows, stride 0.5) -> PN2SSG fused forward per chunk (batch 1, variable size, eager) -> '
                                         'device vote accumulation -> labels; wall clock incl. host launch overhead'}
    return out


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--chunks-per-gpu', type=int, default=32)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-cuda-graph', action='store_true', help='time eager launches instead of one CUDA-graph replay per step')
    ap.add_argument('--no-extras', action='store_true', help='skip e2e / stage / reference-kernel passes (profiling runs)')
    ap.add_argument('--tf32-2d', action='store_true', help='allow TF32 in the cuDNN 2D network (breaks the 1e-4 logit parity; reported in config)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    config = {'workload': 'BASELINE config 3/4: full MVPNet forward, %d chunks per GPU per step (8192 pts, 5 views 160x120), '
                          'unproject + 2D->3D 3-NN + UNetResNet34 + FeatureAggregation + PN2SSG%s' %
                          (args.chunks_per_gpu, ' + NCCL all-gather of logits' if world > 1 else ''),
              'chunks_per_gpu': args.chunks_per_gpu, 'global_chunks': args.chunks_per_gpu * world,
              'parallelism': 'chunk-sharded x%d, replicated weights' % world,
              'l2_policy': 'no flush: per-step working set (~0.8 GB of images/feature maps per GPU) exceeds the 126 MB L2',
              'input_formats': 'as stored by the dataset: colour uint8 HWC, depth uint16 mm, decoded on the device inside every step (both value and e2e)',
              'net_2d_math': 'tcgen05 convolutions of this package, bf16 hi/lo x 3 products, fp32 accumulate (MVPNET_B200_NET2D=cudnn: cuDNN fp32)'}
    DTYPE = 'f32 I/O; contractions as bf16 hi/lo split x 3 tcgen05 products, fp32 accumulate (logits within 1e-4 of the fp32 reference)'

    if args.impl == 'reference':
        if rank != 0:
            return
        res = cpu_arm(max(args.steps, 1), max(args.warmup, 1) if args.warmup else 0)
        line = {'impl': 'reference', 'metric': METRIC, 'value': res['value'], 'unit': 'chunks/s', 'n_gpus': args.gpus,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': res['ms_per_chunk'], 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': DTYPE, 'data': 'synthetic', 'config': config,
                'cpu_baseline': res, 'e2e': {'value': res['value'], 'unit': 'chunks/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line), flush=True)
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the hot path has no CPU fallback); use --impl reference for the CPU arm')
    numa_node = bind_to_gpu_numa(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    torch.backends.cudnn.allow_tf32 = bool(args.tf32_2d)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    import mvpnet_b200
    mvpnet_b200.load_ext()
    from mvpnet_b200 import engine

    model = build_model(device)
    if os.environ.get('MVPNET_B200_UNET_CHANNELS_LAST'):       # strict-fp32 cuDNN is ~1.8x slower in NHWC on B200 (tools/unet_variants.py)
        model.net_2d.to(memory_format=torch.channels_last)
    cpg = args.chunks_per_gpu
    host, _ = make_host_batch([rank * cpg + i for i in range(cpg)], pin=True)

    def to_device(h):
        d = {k: v.to(device, non_blocking=True) for k, v in h.items()}
        d['points_cm'] = d['points'].transpose(1, 2).contiguous()     # (b, 3, np): the reference's `points`
        return d

    from mvpnet_b200.distributed import all_gather_chunks
    gathered = [torch.empty(world * cpg, NUM_CLASSES, NUM_POINTS, device=device) for _ in range(4)] if world > 1 else None
    host_out = torch.empty(cpg, NUM_CLASSES, NUM_POINTS).pin_memory()

    graphed = {'fn': None, 'note': 'eager launches'}

    def batch_of(dev):
        return {'images_u8': dev['images_u8'], 'points': dev['points_cm'], 'depth_mm': dev['depth_mm'], 'pose': dev['pose'],
                'cam_inv': dev['cam_inv'], 'chunk_box': dev['chunk_box'], 'k': KNN}

    def step_device(dev, lane=0):
        with torch.no_grad():
            fn = lane_graphs[lane] if lane_graphs else graphed['fn']
            logit = fn(batch_of(dev)) if fn is not None else hot_path(model, dev)
            if world > 1:
                all_gather_chunks(logit, world * cpg, out=gathered[lane])
        return logit

    step_no = {'i': 0}

    def step_device_laned(dev):
        """One step on the next lane's stream (device-resident inputs)."""
        lane = step_no['i'] % lanes
        step_no['i'] += 1
        if lanes == 1:
            return step_device(dev)
        with torch.cuda.stream(lane_streams[lane]):
            return step_device(dev, lane)

    def lanes_begin():
        for st in lane_streams[:lanes] if lanes > 1 else []:
            st.wait_stream(torch.cuda.current_stream())

    def lanes_drain():
        for st in lane_streams[:lanes] if lanes > 1 else []:
            torch.cuda.current_stream().wait_stream(st)

    def prepare(d):            # device-side view of a freshly uploaded host batch (what to_device does)
        dd = dict(d)
        dd['points_cm'] = d['points'].transpose(1, 2).contiguous()
        return dd

    # end to end through the package's throughput API (engine.PipelinedForward): per step one H2D of the pinned host
    # inputs and one D2H of the logits, on copy streams, double-buffered so that they overlap the neighbouring steps
    pipe = {}
    last = {}
    e2e_mode = {'name': ''}

    submit_s = []

    def step_e2e():
        t0 = time.perf_counter()
        last['out'], last['ev'] = pipe['p'].submit(host)
        submit_s.append(time.perf_counter() - t0)

    def step_e2e_serial():     # the same copies on the compute stream, one after the other
        logit = step_device(to_device(host))
        host_out.copy_(logit, non_blocking=True)
        last['out'] = host_out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, drain=None, begin=None):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if begin is not None:       # other streams start after the start event
            begin()
        for _ in range(n):
            fn()
        if drain is not None:       # work left on other streams (the last download) belongs to the timed region
            drain()
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    dev = to_device(host)
    # Lanes: consecutive steps alternate between `lanes` independent CUDA graphs on their own streams, so that the
    # under-occupied tail of one step (small set-abstraction / propagation levels, 30-150 CTAs) runs beside the
    # convolutions of the next one; the persistent convolution kernels share SMs through their dynamic work distribution.
    lanes = min(4, max(1, int(os.environ.get('MVPNET_B200_BENCH_LANES', '2'))))
    lane_graphs, lane_streams = [], []
    if not args.no_cuda_graph and not args.no_extras:
        try:
            execs = min(4, max(1, int(os.environ.get('MVPNET_B200_BENCH_EXECS', '1'))))

            class RotatingGraph:
                """`execs` instantiations of the same captured forward used in turn: the host can enqueue the next launch
                of a lane while the previous one (another executable graph) is still running."""

                def __init__(self, graphs):
                    self.graphs, self.n, self.cur = graphs, 0, graphs[0]

                def load(self, batch):
                    self.cur = self.graphs[self.n % len(self.graphs)]
                    self.n += 1
                    self.cur.load(batch)

                def replay(self):
                    return self.cur.replay()

                def __call__(self, batch):
                    self.load(batch)
                    return self.replay()

            def new_lane():
                gs = [engine.GraphedForward(model, batch_of(dev)) for _ in range(execs)]
                return gs[0] if execs == 1 else RotatingGraph(gs)

            graphed['fn'] = new_lane()
            graphed['note'] = 'one CUDA graph replay per step (2D network + both streams captured)'
            lane_graphs = [graphed['fn']] + [new_lane() for _ in range(lanes - 1)]
            if execs > 1:
                graphed['note'] += '; %d executable graphs per lane used in turn' % execs
            lane_streams = [torch.cuda.Stream(device) for _ in range(lanes)]
            if lanes > 1:
                graphed['note'] += '; consecutive steps alternate between %d graphs on %d streams' % (lanes, lanes)
        except Exception as ex:      # capture is an optimisation; never lose the measurement over it
            graphed['fn'] = None
            lane_graphs, lane_streams = [], []
            graphed['note'] = 'eager launches (graph capture failed: %s)' % str(ex)[:120]
            torch.cuda.synchronize()
    lanes = len(lane_graphs) if len(lane_graphs) > 1 else 1
    lanes_begin()
    for _ in range(warmup):
        step_device_laned(dev)
    lanes_drain()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.nvtx.range_push('timed')          # ncu --nvtx --nvtx-include "timed/" captures exactly the timed steps
    ms_total = timed(lambda: step_device_laned(dev), args.steps, drain=lanes_drain, begin=lanes_begin)
    torch.cuda.nvtx.range_pop()
    clocks = sampler.summary() if rank == 0 else None
    if args.no_extras:
        if rank == 0:
            print(json.dumps({'metric': METRIC, 'value': cpg * world * args.steps / (ms_total / 1e3), 'unit': 'chunks/s',
                              'note': 'profiling run (--no-extras): not a bench value'}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    try:
        e2e_mode['name'] = 'pipelined (engine.PipelinedForward: H2D / compute / D2H on separate streams, two buffer sets%s)' % (
            ', one compute lane per buffer set' if lanes > 1 else '')
        class LaneForward:
            """Two-phase forward of one lane for engine.PipelinedForward: load() copies the uploaded batch into the lane's
            captured static inputs (the upload buffers are then free), replay() runs the graph (+ the all-gather at N > 1)."""

            def __init__(self, lane):
                self.lane = lane
                self.g = lane_graphs[lane] if lane_graphs else graphed['fn']

            def load(self, d):
                self.g.load(batch_of(d))

            def replay(self):
                with torch.no_grad():
                    logit = self.g.replay()
                    if world > 1:
                        all_gather_chunks(logit, world * cpg, out=gathered[self.lane])
                return logit

        if lane_graphs or graphed['fn'] is not None:
            fwd = [LaneForward(l) for l in range(lanes)] if lanes > 1 else LaneForward(0)
        else:
            fwd = [(lambda d, l=l: step_device(d, l)) for l in range(lanes)] if lanes > 1 else step_device
        pipe['p'] = engine.PipelinedForward(fwd, host, device, prepare=prepare, depth=2)
        for _ in range(2):
            step_e2e()
        del submit_s[:]
        ms_e2e = timed(step_e2e, args.steps, drain=lambda: torch.cuda.current_stream().wait_event(last['ev']))
        e2e_mode['host_submit_ms'] = {'mean': round(1e3 * float(np.mean(submit_s)), 3), 'max': round(1e3 * float(np.max(submit_s)), 3)}
        # the pipelined path returns what the direct call returns (same inputs every step)
        if not torch.allclose(last['out'], step_device(dev).cpu(), rtol=0, atol=1e-5):
            raise RuntimeError('pipelined result differs from the direct forward')
    except Exception as ex:      # never lose the end-to-end number over the pipelining
        torch.cuda.synchronize()
        e2e_mode['name'] = 'serial on the compute stream (pipelined path failed: %s)' % str(ex)[:120]
        for _ in range(2):
            step_e2e_serial()
        ms_e2e = timed(step_e2e_serial, args.steps)

    # per-stage device time of this package's kernels (events on the launching stream), no overlap
    stage_ms = {}
    with torch.no_grad(), engine.profile() as prof:
        reps = max(3, min(args.steps, 5))
        for _ in range(reps):
            hot_path(model, dev, overlap=False)
        for k, v in prof.summary().items():          # stages launched several times per forward: per-forward sums
            per_fwd = np.asarray(v, dtype=np.float64).reshape(reps, -1).sum(axis=1)
            stage_ms[k] = float(np.median(per_fwd))

    gather_check = None
    if world > 1:
        # the all-gathered logits must equal what a single GPU computes for the same chunks: rank 0 recomputes rank 1's shard
        torch.cuda.synchronize()
        step_device(dev, 0)
        torch.cuda.synchronize()
        if rank == 0:
            other, _ = make_host_batch([1 * cpg + i for i in range(cpg)], pin=False)
            with torch.no_grad():
                mine1 = hot_path(model, to_device(other))
            diff = float((gathered[0][cpg:2 * cpg] - mine1).abs().max() / mine1.abs().max())
            gather_check = {'rank1_shard_vs_local_recompute_rel_err': diff, 'ok': diff < 1e-5}
            if diff >= 1e-5:
                raise RuntimeError('bench: all-gathered logits of rank 1 differ from a local recomputation (%.3g)' % diff)
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    try:
        with torch.no_grad():
            launches_mine, launches_all = count_kernel_launches(lambda: hot_path(model, dev))
    except Exception:           # CUPTI unavailable: fall back to the count of profiles/r1_launches_*
        launches_mine, launches_all = KERNELS_PER_STEP, KERNELS_PER_STEP + 11
    traffic_mb, traffic_src = ncu_traffic_mb()
    chunks = cpg * world * args.steps
    value = chunks / (ms_total / 1e3)
    e2e_value = chunks / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = host_out.numel() * host_out.element_size()
    bytes_pc = algorithmic_bytes_per_chunk()
    flops_pc = algorithmic_gflop_per_chunk()
    peak, peak_src = peaks()
    tpeak = tensor_peak()
    mine = {k: v for k, v in stage_ms.items() if k != 'net_2d'}          # net_2d is the sum of its net_2d/* stages
    stages = {}
    for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1]):
        e = {'ms': round(v, 4), 'bound': stage_bound(k)}
        if k in bytes_pc:
            e['alg_MB'] = round(bytes_pc[k] * cpg / 1e6, 3)
            e['GBps'] = round(bytes_pc[k] * cpg / (v / 1e3) / 1e9, 1)
            e['hbm_frac'] = round(e['GBps'] / peak, 4)
        if k in flops_pc:
            e['alg_GFLOP'] = round(flops_pc[k] * cpg, 2)
            e['TFLOPs_alg'] = round(flops_pc[k] * cpg / v, 1)                 # fp32-equivalent useful flops
            e['TFLOPs_issued_bf16'] = round(3 * flops_pc[k] * cpg / v, 1)     # 3 bf16 MMAs per useful MAC (hi/lo split)
            e['tensor_frac_issued'] = round(3 * flops_pc[k] * cpg / v / tpeak, 4)
        stages[k] = e
    # headline roofline: the dominant kernel of this package that is HBM- or tensor-bound by design (the k-NN searches
    # and FPS are CUDA-core-compute / serial-latency bound: their entries in `stages` carry times and bounds instead)
    fused = [k for k in mine if k in flops_pc and not k.startswith('net_2d')]
    # headline roofline: the kernel family of this package that takes the most device time in the step
    launches = {'net_2d/conv3x3': 33, 'net_2d/conv_general': 11}
    top = max([k for k in mine if k in flops_pc and k != 'net_2d'], key=mine.get)
    t_fused = sum(mine[k] for k in fused)
    gf_fused = sum(flops_pc[k] for k in fused) * cpg
    n_launch = launches.get(top, 1)
    roof = {'kernel': top + (' (tc_conv3x3_pair_kernel [cta_group::2, every level above 8 image rows] + tc_conv3x3_kernel [the 8 x 10 level], %d launches per step: every 3x3/stride-1 convolution of the UNet)' % n_launch if top == 'net_2d/conv3x3' else ''),
            'bound': 'hbm' if top == 'feature_aggregation' else 'tensor', 'peak_source': peak_src,
            'launches_per_step': n_launch, 'ms_per_launch': mine[top] / n_launch,
            'traffic': traffic_mb.get(top, NCU_TRAFFIC_MB.get(top)),
            'traffic_note': 'MB per launch (mean over the launches of the step), dram__bytes_read + dram__bytes_write of an ncu --set full capture of this '
                            'command, read at run time from %s' % (traffic_src or 'the round-1 constant (profiles/r1_ncu_full_final.md)'),
            'note': 'algorithmic flops per launch = per-chunk figure x %d chunks per step / launches per step; time = CUDA events around every launch, summed per step' % cpg}
    if roof['bound'] == 'hbm':
        ach = bytes_pc[top] * cpg / (mine[top] / 1e3) / 1e9
        roof.update({'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak})
    else:
        ach = 3 * flops_pc[top] * cpg / mine[top]
        sustained = tensor_peak_sustained()
        roof.update({'achieved': ach, 'peak': sustained, 'unit': 'TFLOP/s', 'frac': ach / sustained,
                     'peak_kind': 'bf16_tflops_sustained of MEASURED_PEAKS.json (kernels timed inside a long step); burst peak %.1f -> frac %.4f' % (tpeak, ach / tpeak),
                     'note2': 'issued bf16 tensor flops (3 products per useful MAC); useful fp32-equivalent = achieved / 3'})
    bq_group = mine.get('ball_query1', 0) + mine.get('set_abstraction1', 0)
    line = {'metric': METRIC, 'value': value, 'unit': 'chunks/s', 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': DTYPE, 'data': 'synthetic', 'config': config, 'launch_mode': graphed['note'], 'numa_node': numa_node, 'numa_note': NUMA_NOTE, 'host_cpus': len(os.sched_getaffinity(0)), 'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'chunks/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / args.steps, 'mode': e2e_mode['name'], 'host_submit_ms': e2e_mode.get('host_submit_ms')},
            'gpu_launches': launches_mine * args.steps,
            'gpu_launches_note': '%d kernels of this package + %d ATen kernels per step, counted from CUPTI activity records of one eager step '
                                 '(torch.profiler); the timed steps replay the same launches from CUDA graphs' % (launches_mine, launches_all - launches_mine),
            'roofline': roof,
            'fused_mlp_family': {'kernels': len(fused), 'ms_per_step': round(t_fused, 4), 'alg_GFLOP_per_step': round(gf_fused, 1),
                                 'TFLOPs_alg': round(gf_fused / t_fused, 1), 'TFLOPs_issued_bf16': round(3 * gf_fused / t_fused, 1),
                                 'tensor_frac_issued': round(3 * gf_fused / t_fused / tpeak, 4), 'peak_TFLOPs_bf16': tpeak},
            'north_star_targets': {
                'ball_query+group SA1 (20.93 MB/chunk unfused-algorithmic, fused here)': {
                    'ms': round(bq_group, 4), 'GBps': round(20.93e6 * cpg / (bq_group / 1e3) / 1e9, 1) if bq_group else None,
                    'hbm_frac': round(20.93e6 * cpg / (bq_group / 1e3) / 1e9 / peak, 4) if bq_group else None}},
            'stages': stages,
            'hot_path_ms_per_step_excl_net2d': sum(v for k, v in mine.items() if not k.startswith('net_2d'))}
    if world == 1:
        try:
            rk = reference_kernel_times(dev['points'])
        except Exception as ex:  # the cross-check must never break the benchmark line
            rk = {'error': str(ex)[:200]}
        if rk is not None:
            line['reference_cuda_kernels_same_gpu'] = rk
    if gather_check is not None:
        line['all_gather_check'] = gather_check
    if world == 1:
        try:
            line['sub_lines'] = bench_extras(model, device, dev, cpg)
        except RuntimeError:
            raise
        except Exception as ex:      # an extra must never cost the headline line
            line['sub_lines'] = {'error': str(ex)[:300]}
    if not args.no_cpu_baseline and world == 1:
        line['cpu_baseline'] = cpu_arm(steps=8, warmup=1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
