import json,sys
d=json.load(open(sys.argv[1]))
w=d.pop('workloads',{})
print('C2 value %.4g'%d['value'], {a:round(b,1) for a,b in d['phase_ms_per_step'].items()}, 'e2e %.4g'%d['e2e']['value'], 'roof %.3f'%d['roofline']['frac'])
for k,v in w.items():
    print(k, 'value %.4g %s'%(v['value'],v['unit']), 'ms/step %.1f'%v['ms_per_step'], 'e2e %.4g'%v['e2e']['value'], 'd2h GB/s %.1f'%v['e2e']['d2h_gbs_per_gpu'], 'roof %.3f'%v['roofline']['frac'], 'emit %.3f'%v['roofline_emit']['frac'])
    print('   ', {a:round(b,1) for a,b in v['phase_ms_per_step'].items()})
