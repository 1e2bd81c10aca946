import sys; sys.path.insert(0,'/root/repo')
import torch, bench, graph_neural_net_b200 as pkg
from graph_neural_net_b200.training import train_step_flat, FlatAdam
from oracle import fgnn_oracle as O
cfg=bench.WORKLOADS["cfg2_er_n200_c32_b128_fwd"]
node_emb=dict(type="node_embedding",block_init="block_emb",block_inside="block",num_blocks=4,in_features=32,out_features=32,depth_of_mlp=3)
m=pkg.models.Siamese_Node_Exp(2,node_emb); m.load_state_dict(bench.make_state_dict(cfg)); m=m.cuda().set_precision("fp16")
x1,x2=bench.make_inputs(cfg,32,1); x1=x1.cuda(); x2=x2.cuda()
opt=FlatAdam(m.parameters(), lr=1e-3)
for _ in range(3): print(train_step_flat(m,opt,{"input":x1},{"input":x2}))
torch.cuda.synchronize()
