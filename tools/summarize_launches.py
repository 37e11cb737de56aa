import csv, collections, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for row in csv.DictReader(lines):
    name = row['Kernel Name']; v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    if unit == 'ns': v /= 1e6
    elif unit == 'us': v /= 1e3
    k = re.sub(r'\(.*', '', name); k = re.sub(r'<.*', '', k)[:90]
    if 'at::' in name:
        m2 = re.search(r'(\w+Functor|\w+_kernel_cuda|GeluCUDAKernelImpl|CatArrayBatchedCopy\w*|layer_norm\w*|RowwiseMoments\w*|GroupNorm\w*|upsample\w*)', name)
        k = 'at:: ' + (m2.group(1) if m2 else name[10:60])
    agg[k][0] += 1; agg[k][1] += v; tot += v
print('total ms', round(tot, 2), 'launches', sum(a[0] for a in agg.values()))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f'{t:8.3f} ms {n:5d}  {k}')
