"""Build libcfpp.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.  No torch headers involved."""
import os, subprocess, sys, hashlib, glob
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'libcfpp.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v' if os.environ.get('CFPP_PTXAS_V') else '-O3']


def sources():
    return sorted(glob.glob(os.path.join(HERE, '*.cu')))


def _digest():
    h = hashlib.sha256()
    for p in sources() + sorted(glob.glob(os.path.join(HERE, '*.cuh'))) + [os.path.join(HERE, '..', '..', 'include', 'cfpp.h')]:
        h.update(open(p, 'rb').read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    stamp = os.path.join(HERE, 'build', 'stamp')
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)

    def cc(src):
        obj = os.path.join(HERE, 'build', os.path.basename(src)[:-3] + '.o')
        cmd = [NVCC, *FLAGS, '-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        if verbose and (r.stderr.strip() or r.stdout.strip()):
            print(r.stdout + r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(cc, sources()))
    r = subprocess.run([NVCC, '-shared', '-o', OUT, *objs, '-lcudart'], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    open(stamp, 'w').write(dig)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
