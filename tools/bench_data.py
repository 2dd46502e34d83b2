"""Throughput of the training augmentation pipeline (make_coco_transforms('train'), src/datasets/coco.py:74-103) on 480x640 samples:
the reference's PIL implementation on one host core (what a DataLoader worker does) vs gw-depth_b200/data.py on the GPU."""
import os
import random
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
from PIL import Image  # noqa: E402
import test_data_gpu as TG  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T, C = TG._reference_transforms()
data = TG._data()
args = types.SimpleNamespace(eval=False)
ref_tf, our_tf = C.make_coco_transforms("train", args), data.make_coco_transforms("train", args)
samples = [TG._sample(s) for s in range(8)]


def run_ref(n):
    done = 0
    for i in range(n):
        img, depth, seg, target = samples[i % 8]
        try:
            ref_tf(Image.fromarray(img), {k: v.clone() for k, v in target.items()},
                   aux_mats=[Image.fromarray(depth, mode="I"), Image.fromarray(seg, mode="L")])
            done += 1
        except ImportError:
            pass
    return done


def run_ours(n, dev):
    done = 0
    for i in range(n):
        img, depth, seg, target = dev[i % 8]
        try:
            our_tf(img, {k: v.clone() for k, v in target.items()}, aux_mats=[depth, seg])
            done += 1
        except ImportError:
            pass
    torch.cuda.synchronize()
    return done


torch.set_num_threads(1)
random.seed(0); torch.manual_seed(0)
run_ref(4)
t0 = time.perf_counter(); n_ref = run_ref(N); t_ref = time.perf_counter() - t0
dev = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(c).cuda(), t) for a, b, c, t in samples]
random.seed(0); torch.manual_seed(0)
run_ours(8, dev)
t0 = time.perf_counter(); n_our = run_ours(N, dev); t_our = time.perf_counter() - t0
print("reference (PIL, 1 host core): %.1f samples/s   gw-depth_b200.data (1 GPU, 1 host thread): %.1f samples/s   x%.1f"
      % (n_ref / t_ref, n_our / t_our, (n_our / t_our) / (n_ref / t_ref)))
