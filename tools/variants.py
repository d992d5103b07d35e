"""Development aid: build kernel variants side by side and time them on the GPU box.

    python tools/variants.py build name1:-DFOO=1,-DBAR name2:...   (here; nvcc cross-compiles)
    python tools/variants.py run [c2|c3:NCOL ...]                   (under gpurun)

Variants are fast builds (NCH = 3, bezier3 only) of the same C-ABI library, written to
build/variants/<name>.so; `run` loads each through LWB200_LIB, checks parity of a small
problem against the C oracle and prints the formal-solution kernel time."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, 'build', 'variants')
SRC = os.path.join(ROOT, 'lightweaver_b200', 'csrc', 'lwb200_api.cu')


def build(specs):
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition(':')
        out = os.path.join(VDIR, name + '.so')
        cmd = ['nvcc', '-O3', '-std=c++17', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a',
               '-Xcompiler', '-fPIC', '-shared', '-DLWB200_DEV_FAST_BUILD', '-Xptxas', '-v',
               *[f for f in flags.split(',') if f], '-o', out, SRC, '-lcudart']
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, pr in procs:
        out, _ = pr.communicate()
        lines = out.splitlines()
        for i, ln in enumerate(lines):
            if ('ray_kernel' in ln or 'gamma_kernel' in ln) and 'Compiling' in ln:
                print(name, ln.split("'")[1][:60], '|', lines[i + 1].strip(), '|', lines[i + 2].strip())
        if pr.returncode:
            print(name, 'FAILED\n', out[-3000:])


def run(workloads):
    for f in sorted(os.listdir(VDIR)):
        if not f.endswith('.so'):
            continue
        env = dict(os.environ, LWB200_LIB=os.path.join(VDIR, f))
        first = True
        for w in workloads:
            wl, _, ncol = w.partition(':')
            cmd = [sys.executable, os.path.join(ROOT, 'tools', 'prof_c3.py'), ncol or '1',
                   '12' if wl != 'c3' else '4', wl] + (['check'] if first else [])
            first = False
            r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
            for ln in (r.stdout + (r.stderr[-1500:] if r.returncode else '')).splitlines():
                print(f'[{f[:-3]}]', ln, flush=True)


if __name__ == '__main__':
    if sys.argv[1] == 'build':
        build(sys.argv[2:])
    else:
        run(sys.argv[2:] or ['c2', 'c3:256'])
