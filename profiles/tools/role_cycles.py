import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cmrl_b200 as cm
sys.argv=['bench.py']
import bench
SILU=bench.SILU
dev='cuda:0'
torch.manual_seed(0); np.random.seed(0)
N=100000
tr = cm.ExternalMaskTransition(5,1,hid_size=200,activation_fn_cfg=SILU,device=dev)
tr.set_input_mask(torch.from_numpy(bench.make_mask()))
tr.set_elite_members([0,1,2,3,4])
dyn = cm.PlainEnsembleDynamics(tr, learned_reward=False)
env = cm.VecFakeEnv(N,None,None)
rng=np.random.default_rng(0)
env.set_up(dyn, reward_fn=lambda n,o,a: np.zeros((n.shape[0],1),np.float32), termination_fn=lambda n,o,a: np.zeros((n.shape[0],1),bool), get_init_obs_fn=lambda n: rng.normal(size=(n,5)).astype(np.float32))
env.precision='f16'; env.reset()
act=torch.rand(N,1,device=dev)*2-1
idx,noise=cm.ops.draw_member_noise(1,1,N,env._elite_u8('transition'),5)
for _ in range(3): env.predict_selected(tr, env._obs_dev, act, idx, noise)
lib=cm._lib.load(); lib.cmrl_debug_set_buffer.argtypes=[ctypes.c_void_p]
for flags in [0]:
  dbg=torch.zeros(148*16,dtype=torch.int64,device=dev)
  lib.cmrl_debug_set_flags(flags); lib.cmrl_debug_set_buffer(dbg.data_ptr())
  env.predict_selected(tr, env._obs_dev, act, idx, noise); torch.cuda.synchronize()
  lib.cmrl_debug_set_buffer(None); lib.cmrl_debug_set_flags(0)
  d=dbg.cpu().numpy().reshape(148,16)
  lead=d[0::2]
  print('=== variant', cm.ops.TC_VARIANT, 'dbg flags', flags, '(1 skip A stores, 4 skip activation)')
  print('MMA thread (leader CTAs): total, wait_x, wait_a, issue, items  (mean over clusters)')
  print(lead[:,:5].mean(0).astype(int), ' per item:', (lead[:,:4].mean(0)/lead[:,4].mean()).astype(int))
  print('epilogue warp0: total, wait_d, loop, wait_h, head, ld_wait, publish')
  print(d[:,8:15].mean(0).astype(int), ' per item:', (d[:,8:15].mean(0)/lead[:,4].mean()).astype(int))
