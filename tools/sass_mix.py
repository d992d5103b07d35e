"""Development aid (runs here, no GPU): static SASS opcode mix of one kernel of a built library.

    python tools/sass_mix.py <lib.so> <kernel name substring> [more substrings that must also match]

Prints registers / spills from the ELF notes if available and a histogram of opcodes; `--dump` writes the
kernel's SASS to stdout instead."""
import collections
import re
import subprocess
import sys


def kernels(lib):
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    cur, name = [], None
    for ln in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', ln)
        if m:
            if name:
                yield name, cur
            name, cur = m.group(1), []
        elif name:
            cur.append(ln)
    if name:
        yield name, cur


def main():
    lib, subs = sys.argv[1], [a for a in sys.argv[2:] if not a.startswith('--')]
    dump = '--dump' in sys.argv
    for name, lines in kernels(lib):
        if not all(s in name for s in subs):
            continue
        if dump:
            print('\n'.join(lines))
            return
        ops = collections.Counter()
        for ln in lines:
            m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)', ln)
            if m:
                ops[m.group(1)] += 1
        tot = sum(ops.values())
        fp64 = sum(v for k, v in ops.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
        print(f'{name}\n  {tot} instructions, fp64 {fp64} ({100.0 * fp64 / tot:.0f} %)')
        print('  ' + '  '.join(f'{k} {v}' for k, v in ops.most_common(24)))


if __name__ == '__main__':
    main()
