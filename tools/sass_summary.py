"""Per-kernel counts of the SASS mnemonics that prove which hardware paths a kernel uses:
   python tools/sass_summary.py [mv2d_b200/lib/libmv2d_b200.so] > profiles/rNN_sass_summary.txt
UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = TMA tensor load / store,
UBLKCP = cp.async.bulk, SYNCS = mbarrier, HMMA = warp-level mma.sync (TF32), FFMA2 = packed fp32 FMA."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else 'mv2d_b200/lib/libmv2d_b200.so'
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
keys = ['UTCHMMA', 'UTCQMMA', 'UTCMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'HMMA', 'FFMA2', 'FFMA', 'DFMA']
counts = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace('(anonymous namespace)::', '').replace('mv2d::', '')
        name = re.sub(r'\(.*', '', name)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m:
        op = m.group(1)
        for k in keys:
            if op == k or (k in ('UTCHMMA', 'UTCQMMA', 'UTCMMA') and op.startswith(k)):
                counts[name][k] += 1
                break
arch = re.search(r'arch = (sm_\w+)', out)
print(f'# {so}: SASS mnemonic counts per kernel ({arch.group(1) if arch else "?"}); kernels without any of the tensor / TMA / bulk-copy opcodes are listed at the end')
print(f'{"kernel":70s} ' + ' '.join(f'{k:>8s}' for k in keys))
plain = []
for n, c in counts.items():
    if not any(c[k] for k in keys[:9]):
        plain.append(n)
        continue
    print(f'{n[:70]:70s} ' + ' '.join(f'{c[k]:8d}' for k in keys))
print('\n# FFMA / FFMA2 kernels (no tensor-core, TMA or bulk-copy opcodes):')
for n in plain:
    c = counts[n]
    print(f'{n[:70]:70s} ' + ' '.join(f'{c[k]:8d}' for k in keys))
