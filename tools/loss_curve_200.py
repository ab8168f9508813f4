import sys, time, json
sys.path.insert(0, '.')
import torch, numpy as np
from oracle.golden_inputs import train_batch
from oracle.init_state import init_state
from oracle.train_step import AdamW as OracleAdamW, split_state, train_step
from torch_semantic_segmentation_b200.engine import create_segmentation_trainer, Events
from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
from torch_semantic_segmentation_b200.models import fastscnn
x, y = train_batch('fastscnn')
torch.set_num_threads(16)
t0=time.time()
sd = split_state(init_state('fastscnn', 0)); oopt = OracleAdamW(sd, lr=1e-3, weight_decay=1e-5)
ref = [train_step('fastscnn', sd, oopt, x, y, dropout_mask=1.0) for _ in range(200)]
print('oracle s', time.time()-t0)
res={'ref':ref}
for name, dtype in (('fp32', torch.float32), ('bf16', torch.bfloat16)):
    torch.manual_seed(0)
    model = fastscnn(3, 19).cuda().set_compute_dtype(dtype)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout): m.p = 0.0
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    tr = create_segmentation_trainer(model, opt, CrossEntropyLoss(ignore_index=255), 'cuda', logging=False)
    losses=[]
    tr.add_event_handler(Events.ITERATION_COMPLETED, lambda e: losses.append(e.state.output))
    tr.run([(x, y)] * 200)
    a=np.array(losses); r=np.array(ref)
    relerr=np.abs(a-r)/r
    print(name, 'first', a[0], r[0], 'last', a[-1], r[-1], 'max rel', relerr.max(), 'mean rel', relerr.mean(), 'rel@[10,50,100,199]', relerr[[10,50,100,199]])
    res[name]=losses
json.dump(res, open('gpurun_out/loss_curves_200.json','w'))
