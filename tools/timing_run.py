import sys; sys.path.insert(0,'/root/repo')
import torch, bench, graph_neural_net_b200 as pkg
cfg=bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
node_emb=dict(type="node_embedding",block_init="block_emb",block_inside="block",num_blocks=4,in_features=64,out_features=64,depth_of_mlp=3)
m=pkg.models.Siamese_Node_Exp(2,node_emb); m.load_state_dict(bench.make_state_dict(cfg)); m=m.cuda().set_precision("fp16")
x1,x2=bench.make_inputs(cfg,8,1); x1=x1.cuda()
lib=pkg.get_lib()
with torch.no_grad():
    for _ in range(2): m.embed({"input":x1})
    torch.cuda.synchronize(); lib.fgnn_debug_dump_timing()
    m.embed({"input":x1}); torch.cuda.synchronize()
lib.fgnn_debug_dump_timing()
