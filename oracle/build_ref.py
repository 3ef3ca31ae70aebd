#!/usr/bin/env python
"""Vendor the reference's own implementation of the hot path into oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST / BENCH INFRASTRUCTURE.  The reference is pure Python: "building" it means copying the three files of the path
(models/satrans.py, models/meta_basemodel.py, models/submodules.py -- plus the package __init__) from /root/reference,
unmodified, into oracle/_ref/models/ so that they travel to the GPU box with the snapshot (where /root/reference does
not exist).  Nothing is ever copied into tracked directories.  `bench.py --impl reference` imports them through
oracle/ref_shims (name plumbing for the un-installed deepctr_torch / tensorflow / torch_scatter) and then reports
`kind: "reference"`; without oracle/_ref it falls back to the oracle port (`kind: "port"`).

    python oracle/build_ref.py            # also run by __graft_entry__.build() when /root/reference is present
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('SATRANS_REFERENCE', '/root/reference')
FILES = ['models/__init__.py', 'models/satrans.py', 'models/meta_basemodel.py', 'models/submodules.py']


def build(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print('oracle/build_ref: %s not present, nothing to do' % REF)
        return False
    dst_root = os.path.join(HERE, '_ref')
    for rel in FILES:
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    if verbose:
        print('oracle/build_ref: %d reference files -> %s' % (len(FILES), dst_root))
    return True


def available():
    return all(os.path.exists(os.path.join(HERE, '_ref', rel)) for rel in FILES)


def import_reference():
    """(SATrans, SparseFeat, DenseFeat) of the vendored, unmodified reference."""
    for p in (os.path.join(HERE, 'ref_shims'), os.path.join(HERE, '_ref')):
        if p not in sys.path:
            sys.path.insert(0, p)
    from models.satrans import SATrans
    from deepctr_torch.inputs import DenseFeat, SparseFeat
    return SATrans, SparseFeat, DenseFeat


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
