import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['config']['name'], 'its/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'roofline_iter', round(d['roofline_iter']['frac'], 3), 'clocks', d['clocks'])
for k in d.get('kernels') or []:
    print('   ', k['name'][:46], round(k['avg_us'], 1), 'us', round(k['frac_of_peak'], 3), 'share', round(k['share'], 3))
for n, e in (d.get('extra') or {}).items():
    if 'error' in e:
        print(n, e); continue
    print(n, 'its/s', round(e['its_per_s'], 1), 'ms/step', round(e['ms_per_step'], 3), 'roofline_iter', round(e['roofline_iter']['frac'], 3))
    for k in e.get('kernels', []):
        print('   ', k['name'][:46], round(k['avg_us'], 1), 'us', round(k['frac_of_peak'], 3), 'share', round(k['share'], 3))
print('e2e', d.get('e2e'))
print('cpu', d.get('cpu_baseline'))
