"""
make_config1.py — BASELINE configs[0] ("config 1"): the reference's own CPU-runnable case, run by the REFERENCE ITSELF.

One 1024 x 1024 uint8 tile through the reference's PanopticDeepLab (ResNet-50 encoder, the MitoNet configuration of
configs/mmm_panoptic_deeplab_pointrend.yaml:8-20, torch.manual_seed(0), random-initialised, eval mode, CPU) and its
PanopticDeepLabEngine (empanada/inference/engines.py:92-160).  A random-initialised network emits per-channel constants
(SURVEY appendix A6), so seeded synthetic head tensors (~60 ellipses) are added to its outputs, as SURVEY 8d prescribes,
before the reference's post-processing runs.

Writes
  tests/golden/config1.npz          the network's per-channel constants, the generator's parameters and the
                                    reference's panoptic map (the head tensors are regenerated from the seed by
                                    empanada_b200.synth.config1_heads, which the GPU test and bench.py call too)
  profiles/r2_config1_cpu.json      seconds of the reference's CNN forward and post-processing on this container's CPU

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_config1.py
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from empanada_b200.synth import CONFIG1 as PARAMS, CONFIG1_TILE, config1_heads  # noqa: E402

HW, N_INSTANCES, SEED = CONFIG1_TILE


def main():
    sys.path.insert(0, '/root/reference')
    import torch
    from empanada.inference import engines as reng                      # reference
    from empanada.models.panoptic_deeplab import PanopticDeepLab         # reference
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = PanopticDeepLab(encoder='resnet50', num_classes=1, low_level_stages=[1], low_level_channels_project=[32],
                          ins_decoder=True).eval()
    tile = np.clip(np.random.default_rng(SEED).normal(130, 35, (HW, HW)), 0, 255).astype(np.uint8)
    image = ((torch.from_numpy(tile).float() / 255 - 0.508979) / 0.148561)[None, None]
    with torch.no_grad():
        raw = net(image)
    consts = []
    for k in ('sem_logits', 'ctr_hmp', 'offsets'):
        for ch in range(raw[k].size(1)):
            v = raw[k][0, ch]
            assert float(v.min()) == float(v.max()), f'{k}[{ch}] is not constant'
            consts.append(float(v.flatten()[0]))
    consts = np.asarray(consts, np.float32)
    heads = {k: torch.from_numpy(v) for k, v in config1_heads(consts).items()}

    class WithSyntheticHeads(torch.nn.Module):
        """the reference network with the synthetic head tensors added to what it emits"""

        def __init__(self):
            super().__init__()
            self.net = net

        def forward(self, x):
            out = self.net(x)
            for k in ('sem_logits', 'ctr_hmp', 'offsets'):
                assert out[k].shape == heads[k].shape
            return {k: (heads[k] if k in heads else v) for k, v in out.items()}

    # heads[k] already IS constant + synthetic (the float32 adds are done in numpy, by config1_heads)
    engine = reng.PanopticDeepLabEngine(WithSyntheticHeads(), **PARAMS)
    t_all, t_cnn, t_post = [], [], []
    for _ in range(3):
        t0 = time.perf_counter()
        pan = engine(image)
        t_all.append(time.perf_counter() - t0)
        with torch.no_grad():
            t0 = time.perf_counter()
            out = engine.infer(image)
            t_cnn.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            sem = engine._harden_seg(out['sem'])
            pan2 = engine.postprocess(sem, out['ctr_hmp'], out['offsets'])
            t_post.append(time.perf_counter() - t0)
        assert torch.equal(pan, pan2)
    pan = pan.numpy()
    n_inst = int(np.unique(pan[pan > 0]).size)
    np.savez_compressed(os.path.join(HERE, 'config1.npz'), consts=consts, pan=pan.astype(np.int32),
                        params=json.dumps(dict(PARAMS, hw=HW, n_instances=N_INSTANCES, seed=SEED)))
    rec = {'workload': 'config 1: 1024x1024 tile, PanopticDeepLab ResNet-50 (random init) + synthetic heads, reference code on CPU',
           'threads': threads, 'torch': torch.__version__, 'instances': n_inst,
           'seconds_engine_call': float(np.median(t_all)), 'seconds_cnn_forward': float(np.median(t_cnn)),
           'seconds_postprocess': float(np.median(t_post)),
           'postprocess_mpix_per_s': HW * HW / float(np.median(t_post)) / 1e6,
           'note': 'median of 3; measured in the build container (no GPU); bench.py "config1" times the same tile on the B200 '
                   'and checks its panoptic map against tests/golden/config1.npz'}
    with open(os.path.join(ROOT, 'profiles', 'r2_config1_cpu.json'), 'w') as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))
    print('config1.npz:', os.path.getsize(os.path.join(HERE, 'config1.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
