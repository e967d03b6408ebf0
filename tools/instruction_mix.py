"""Executed-instruction mix of the stream pipeline's kernels from a `ncu --set full --import-source on` capture.
usage: python tools/instruction_mix.py <capture.ncu-rep> <chunks per pass> <out.txt>"""
import collections, csv, re, subprocess, sys
rep, nchunks, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
text = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
starts = [i for i, l in enumerate(text) if l.startswith('"Kernel Name"')] + [len(text)]
seen = set()
sha = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
with open(out, 'w') as f:
    f.write(f'# executed warp instructions per 2 KiB chunk, by opcode (ncu --set full --import-source on, SASS view; 1 GiB synthetic document, git {sha})\n')
    for a, b in zip(starts, starts[1:]):
        full = next(csv.reader([text[a]]))[1].replace('void ', '')
        name = re.sub(r'\((?:int|bool)\)', '', full[:full.index('>(') + 1] if '>(' in full else full.split('(')[0])
        rows = list(csv.reader(text[a + 1:b]))
        hdr = rows[0]
        if 'Instructions Executed' not in hdr or name in seen or 'SASS' in ''.join(hdr[:1]):
            continue
        if not rows[1][hdr.index('Source')].lstrip().split(' ')[0].isupper() and '@' not in rows[1][hdr.index('Source')]:
            continue   # the CUDA-C view of the same kernel
        seen.add(name)
        iI, iS, iT = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('Thread Instructions Executed')
        data = [r for r in rows[1:] if len(r) == len(hdr)]
        tot = sum(int(r[iI]) for r in data)
        thr = sum(int(r[iT]) for r in data)
        ops = collections.Counter()
        for r in data:
            m = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iS])
            if m:
                ops[m.group(1)] += int(r[iI])
        f.write(f'\n{name}: {tot / nchunks:.1f} per chunk, {thr / max(tot, 1):.1f} active threads per instruction\n')
        f.write('  ' + ' '.join(f'{k}:{v / nchunks:.1f}' for k, v in ops.most_common(28)) + '\n')
print(open(out).read())
