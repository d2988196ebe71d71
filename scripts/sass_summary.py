"""Count the Blackwell-specific instructions per kernel family in the built library (no GPU needed):
    python scripts/sass_summary.py > profiles/r2_sass_summary.txt
UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA load), UBLKCP = cp.async.bulk, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, VIADDMNMX = DPX add-min, HMMA = mma.sync, FFMA = fp32 CUDA-core FMA."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'sequitr_b200', 'libsequitr_b200.so')
OPS = ['UTCHMMA', 'UTMALDG', 'UBLKCP', 'LDTM', 'UTCBAR', 'SYNCS', 'VIADDMNMX', 'HMMA', 'FFMA']
FAMILIES = [('conv_qd_kernel', 'conv_qd_kernel'), ('conv_qf_kernel', 'conv_qf_kernel'), ('conv_qu_kernel', 'conv_qu_kernel'),
            ('conv_xc_pair_kernel', 'conv_xc_pair_kernel (opt-in)'), ('conv_xc_kernel', 'conv_xc_kernel'),
            ('conv_tc_kernel', 'conv_tc_kernel'), ('first_conv', 'first_conv_kernel'),
            ('edt_cols_dpx', 'edt_cols_dpx (W1)'), ('inst_cols_dpx', 'inst_cols_dpx (W3)'),
            ('conv_fp32_tile_kernel', 'conv_fp32_tile_kernel (fp32 exact mode, training)'),
            ('wgrad_kernel', 'wgrad_kernel (training)'),
            ('run_|root_emit|ccl_', 'ccl (7 kernels)')]


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    funcs = collections.Counter()
    fam = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = m.group(1)
            fam = 'other'
            for pat, label in FAMILIES:
                if re.search(pat, name):
                    fam = label
                    break
            funcs[fam] += 1
            counts.setdefault(fam, collections.Counter())
            continue
        if fam is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m:
            op = m.group(1)
            if op in OPS:
                counts[fam][op] += 1
    print('# cuobjdump -sass sequitr_b200/libsequitr_b200.so: Blackwell-specific instructions per kernel family '
          '(scripts/sass_summary.py, end of round 2)')
    print('# ' + __doc__.strip().split('\n', 2)[2].replace('\n', ' '))
    print('kernel_family,functions,' + ','.join(OPS))
    for fam in sorted(counts):
        print('%s,%d,%s' % (fam, funcs[fam], ','.join(str(counts[fam][o]) for o in OPS)))


if __name__ == '__main__':
    main()
