import json,sys
d=json.load(open(sys.argv[1]))
for r in d:
    if r['dir']!='wgrad': continue
    print(f"{r['kind']:5s} {r['c_in']:4d}->{r['c_out']:4d} rows {r['n_out']:7d}  wgrad {r['ms']*1000:6.1f} us  frac {r['alg_bytes']/r['ms']/1e6/6534.5:.3f}")
print('wgrad total us', sum(r['ms'] for r in d if r['dir']=='wgrad')*1000)
