"""Builds libampis_b200.so (CUDA kernels + C ABI, sm_100a only) in-tree with nvcc.

    python -m ampis_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repository snapshot.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libampis_b200.so')
DIGEST = LIB + '.digest'

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC,-O3,-Wall,-pthread',
    '--shared',
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')) + glob.glob(os.path.join(CSRC, '*.cpp')))


def source_digest():
    """Digest of everything the library is compiled from (sources, headers, flags).  Content based, not mtime
    based: a snapshot copied to another box keeps its contents but not necessarily its timestamps."""
    import hashlib
    h = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
    deps = sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + \
        [os.path.join(os.path.dirname(HERE), 'include', 'ampis_b200.h')]
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    """True when the library is missing or was compiled from other sources than the ones present (the digest of
    the sources it was built from is kept beside it)."""
    if not os.path.exists(LIB) or not os.path.exists(DIGEST):
        return True
    with open(DIGEST) as f:
        return f.read().strip() != source_digest()


def build(force=False, verbose=False):
    """Compile when the library is missing or older than a source.  Safe when several processes (the ranks of
    ampis_b200.distributed on a fresh checkout) arrive together: an exclusive file lock serialises them, the
    compiler writes to a temporary file that is renamed into place, so nobody ever loads a half-written library."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(LIB + '.lock', 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():          # another process built it while we waited
                return LIB
            nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
            tmp = '%s.tmp.%d' % (LIB, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', tmp] + sources() + ['-lcudart']
            try:
                digest = source_digest()
                subprocess.check_call(cmd)
                os.replace(tmp, LIB)
                with open(DIGEST, 'w') as f:
                    f.write(digest + '\n')
            finally:
                if os.path.exists(tmp):
                    os.unlink(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


def _build_host_lib(src, out, cmd_head, extra_digest=''):
    """Compile one host-only source into `out` when missing or built from other contents (digest beside it);
    locked and renamed into place like build()."""
    import fcntl
    import hashlib
    with open(src, 'rb') as f:
        digest = hashlib.sha256(f.read() + extra_digest.encode()).hexdigest()
    dpath = out + '.digest'
    fresh = lambda: os.path.exists(out) and os.path.exists(dpath) and open(dpath).read().strip() == digest
    if fresh():
        return out
    with open(out + '.lock', 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if fresh():
                return out
            tmp = '%s.tmp.%d' % (out, os.getpid())
            try:
                subprocess.check_call(cmd_head + ['-o', tmp, src])
                os.replace(tmp, out)
                with open(dpath, 'w') as f:
                    f.write(digest + '\n')
            finally:
                if os.path.exists(tmp):
                    os.unlink(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return out


SYNTH_SRC = os.path.join(HERE, 'synth', 'synth.cpp')
SYNTH = os.path.join(HERE, 'libampis_synth.so')


def build_synth():
    """libampis_synth.so: the synthetic micrograph generator (bench / test data only; plain g++, no CUDA)."""
    return _build_host_lib(SYNTH_SRC, SYNTH, [os.environ.get('CXX', 'g++'), '-O3', '-std=c++17', '-shared', '-fPIC',
                                              '-Wall', '-pthread'])


MARSHAL_SRC = os.path.join(HERE, 'cext', 'pymarshal.c')
MARSHAL = os.path.join(HERE, '_pymarshal.so')


def build_marshal(force=False):
    """Compile the CPython extension ampis_b200._pymarshal (host marshaller, plain C, gcc) in-tree."""
    import sysconfig
    if force and os.path.exists(MARSHAL + '.digest'):
        os.unlink(MARSHAL + '.digest')
    return _build_host_lib(MARSHAL_SRC, MARSHAL, [os.environ.get('CC', 'gcc'), '-O2', '-shared', '-fPIC', '-Wall', '-I',
                                                  sysconfig.get_paths()['include']], extra_digest=sys.version)


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    build_marshal(force='--force' in sys.argv)
    build_synth()
    print(LIB)
    print(MARSHAL)
    print(SYNTH)
