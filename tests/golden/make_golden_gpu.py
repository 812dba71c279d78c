"""Generates tests/golden/dfa3d_ref_*.pt on the B200 box from the reference's own kernels (oracle/_ref).

    gpurun -- 'python tests/golden/make_golden_gpu.py'    # writes gpurun_out/golden/*.pt; copy into tests/golden/

Each file holds the seeded inputs' generator arguments and the reference outputs/gradients (fp32), small
enough to commit.  tests/test_oracle_cpu.py checks the CPU oracle against them without a GPU."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import build_ref  # noqa: E402
from test_gpu_ref_ext import make_case, reference_fwd_bwd  # noqa: E402

GOLDEN_CASES = {
    'stage2': ((2, 24, 8, 32, 12, [(7, 10)], 4), 101),
    'stage1': ((2, 24, 1, 256, 12, [(7, 10)], 1), 102),
    'large': ((2, 24, 8, 16, 12, [(7, 10)], 4), 103),
    'multilevel': ((1, 16, 4, 8, 10, [(6, 8), (3, 4)], 3), 104),
}


def main():
    ext = build_ref.load()
    assert ext is not None and torch.cuda.is_available()
    out_dir = os.path.join(ROOT, 'gpurun_out', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    for name, (args, seed) in GOLDEN_CASES.items():
        c = make_case(*args, seed=seed)
        cu = {k: v.cuda() for k, v in c.items()}
        ref = reference_fwd_bwd(ext, cu['value'], cu['dist'], cu['s3'], cu['lsi'], cu['loc'], cu['attn'], cu['gout'])
        blob = dict(args=args, seed=seed, torch=torch.__version__, device=torch.cuda.get_device_name(0),
                    **{k: v.cpu() for k, v in ref.items()})
        torch.save(blob, os.path.join(out_dir, f'dfa3d_ref_{name}.pt'))
        print(name, {k: tuple(v.shape) for k, v in ref.items()})


if __name__ == '__main__':
    main()
