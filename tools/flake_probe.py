import sys, types, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from conftest import load_golden, rel_err
import test_gpu_parity as T
dev=torch.device('cuda:0')
arr,W,m=T.golden_model('small_rsf_refine.npz',dev,True)
T.install_golden_state(m,arr,dev)
g=T.golden_graph(arr,dev)
inp=torch.relu(arr['fct1'][:,64:]).to(dev)
worst={}
for rep in range(60):
    net=torch.tanh(arr['fct1'][:,:64].double()).float().to(dev)
    # churn the allocator a bit
    junk=[torch.randn(1000+rep*37,device=dev) for _ in range(rep%5)]
    with torch.no_grad():
        for it in range(3):
            coords=arr[f'it{it}/coords'].to(dev)
            flow=coords-arr['pc1'].to(dev)
            gcorr=arr[f'it{it}/corr'].to(dev)
            corr=m.corr_block(coords)
            mot=m.update_block.motion_encoder(flow,gcorr)
            net2,delta=m.update_block(net,inp,gcorr,flow,g)
            for name,a,b in (('corr',corr,arr[f'it{it}/corr']),('mot',mot,arr[f'it{it}/motion']),('net',net2,arr[f'it{it}/net']),('delta',delta,arr[f'it{it}/delta'])):
                e=rel_err(a.cpu(),b); k=(name,it)
                worst.setdefault(k,[]).append(e)
            net=arr[f'it{it}/net'].to(dev)
for k,v in sorted(worst.items()):
    print(k, 'min %.2e max %.2e distinct %d'%(min(v),max(v),len(set(v))))
