"""Extract the FAL C 82-point model atmosphere table (Fontenla, Avrett & Loeser
1993, model C) from the reference's data module into a small .npz that ships
with this package.

The reference file (lightweaver/fal.py:8-429) cannot be imported here (it pulls
in astropy through lightweaver/atmosphere.py), so the numeric array literals are
evaluated on their own.  Only numbers are extracted; run once, output committed:

    python tools/make_falc82.py /root/reference lightweaver_b200/data/falc82.npz
"""
import re
import sys

import numpy as np


def main(ref_root, out_path):
    src = open(f'{ref_root}/lightweaver/fal.py').read()
    # keep only the block of array definitions (between the imports and Falc82)
    start = src.index('cmass = ')
    end = src.index('Falc82')
    ns = {'np': np}
    exec(src[start:end], ns)  # array literals only
    out = {k: np.asarray(ns[k], dtype=np.float64) for k in ('cmass', 'temp', 'ne', 'vel', 'vturb', 'nh')}
    assert out['cmass'].shape == (82,) and out['nh'].shape == (6, 82)
    np.savez(out_path, **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
