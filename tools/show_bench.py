import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','stages','cpu_baseline','clocks'):
    print(k, d.get(k))
r=d['roofline']; print('roofline', {k:v for k,v in r.items() if k!='note'})
