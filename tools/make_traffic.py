"""profiles/r2_traffic.json from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list of one steady-state Newton step (tools/r2_ncu_job.sh): DRAM bytes per launch of every kernel family, and the
total of the LDL^T factorisation graph (the launches from ldlt_reset_kernel to ldlt_blockinv_kernel)."""
import collections
import csv
import json
import sys

MUL = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3}


def main(path, out):
    rows = collections.OrderedDict()
    for row in csv.DictReader(l for l in open(path) if not l.startswith('==')):
        r = rows.setdefault(row['ID'], {'name': row['Kernel Name'].split('(')[0].replace('void ', '')})
        r[row['Metric Name']] = float(row['Metric Value'].replace(',', '')) * MUL.get(row['Metric Unit'], 1.0)
    fam = collections.OrderedDict()
    in_factor, fac = False, {'bytes': 0.0, 'us': 0.0, 'launches': 0}
    for r in rows.values():
        b = r.get('dram__bytes_read.sum', 0.0) + r.get('dram__bytes_write.sum', 0.0)
        f = fam.setdefault(r['name'], {'launches': 0, 'bytes': 0.0, 'us': 0.0, 'first': b})
        f['launches'] += 1
        f['bytes'] += b
        f['us'] += r.get('gpu__time_duration.sum', 0.0)
        if r['name'].startswith('ldlt_reset_kernel'):
            in_factor = True
        if in_factor:
            fac['bytes'] += b
            fac['us'] += r.get('gpu__time_duration.sum', 0.0)
            fac['launches'] += 1
        if r['name'].startswith('ldlt_blockinv_kernel'):
            in_factor = False
    res = {'source': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over ONE steady-state Newton '
                     'step at config 3 (tools/r2_ncu_job.sh, tools/make_traffic.py); bytes = dram read + write, per launch '
                     '(named keys: the first launch of the family in the step; per_family: averages); ldlt_factor = all launches of the '
                     'factorisation graph together'}
    for k, f in fam.items():     # named keys: the FIRST launch of the family in the step (residual GEMV, d2L contraction, ...)
        res[k.split('<')[0] if k.split('<')[0] not in res else k] = f['first']
    res['ldlt_factor'] = fac['bytes']
    res['ldlt_factor_launches'] = fac['launches']
    res['per_family'] = {k: {'launches': f['launches'], 'bytes_per_launch': f['bytes'] / f['launches'],
                             'us_per_launch_serialised': f['us'] / f['launches']} for k, f in fam.items()}
    json.dump(res, open(out, 'w'), indent=1)
    print('ldlt_factor: %d launches, %.1f MB' % (fac['launches'], fac['bytes'] / 1e6))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
