#!/usr/bin/env python
"""Mirror of the reference's scripts/train_paac_conv.py (CLI flags :95-121, wiring :30-92):

    python scripts/train_paac_conv.py --height=84 --clip_norm=1                      # config 3: 32 emulators, 1 GPU
    torchrun --nproc-per-node 8 scripts/train_paac_conv.py -ec 8192 --clip_norm=1    # config 5: env batch sharded

Flag names and defaults are the reference's; additions: --n_locusts, --max_updates, --no_cuda_graph,
--reward_indexing, --mask_terminals.  '-d/--device' accepts the reference's '/gpu:0' spelling.
"""
import argparse
import copy
import logging
import os
import signal
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

logging.getLogger().setLevel(logging.INFO)


def bool_arg(string):
    value = string.lower()
    if value == 'true':
        return True
    elif value == 'false':
        return False
    raise argparse.ArgumentTypeError("Expected True or False, but got {}".format(string))


def get_network_and_environment_creator(args, random_seed=3):
    import golds_rl_gym_b200 as pkg
    ec = pkg.submodule("agents.paac.environment_creator")
    pv = pkg.submodule("agents.paac.policy_v_network")
    env_creator = ec.SwarmEnvironmentCreator(n_locusts=args.n_locusts, grid_size=args.height)
    args.num_actions = env_creator.num_actions
    args.random_seed = random_seed
    network_conf = {
        'num_actions': args.num_actions,
        'entropy_regularisation_strength': args.entropy_regularisation_strength,
        'device': args.device, 'height': args.height, 'width': args.height, 'channels': 3,
        'filters': args.filters, 'conv_layers': 2, 'scale': args.scale, 'clip_norm': args.clip_norm,
        'clip_norm_type': args.clip_norm_type, 'static_size': args.static_size,
        'temporal_size': args.temporal_size, 'static_hidden_size': args.static_hidden_size,
        'rnn_hidden_size': args.temporal_hidden_size,
    }

    def network_creator(name='local_learning'):
        conf = copy.copy(network_conf)
        conf['name'] = name
        return pv.ConvSingleAgentPolicyNetwork(conf)

    return network_creator, env_creator


def get_arg_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('-d', '--device', default='/gpu:0', type=str, dest="device")
    parser.add_argument('--e', default=0.1, type=float, dest="e")
    parser.add_argument('--alpha', default=0.99, type=float, dest="alpha")
    parser.add_argument('-lr', '--initial_lr', default=0.0001, type=float, dest="initial_lr")
    parser.add_argument('-lra', '--lr_annealing_steps', default=80000000, type=int, dest="lr_annealing_steps")
    parser.add_argument('--entropy', default=0.02, type=float, dest="entropy_regularisation_strength")
    parser.add_argument('--clip_norm', default=40.0, type=float, dest="clip_norm")
    parser.add_argument('--clip_norm_type', default="global", dest="clip_norm_type")
    parser.add_argument('--gamma', default=0.99, type=float, dest="gamma")
    parser.add_argument('--max_global_steps', default=80000000, type=int, dest="max_global_steps")
    parser.add_argument('--max_local_steps', default=5, type=int, dest="max_local_steps")
    parser.add_argument('--single_life_episodes', default=False, type=bool_arg, dest="single_life_episodes")
    parser.add_argument('-ec', '--emulator_counts', default=32, type=int, dest="emulator_counts")
    parser.add_argument('-ew', '--emulator_workers', default=8, type=int, dest="emulator_workers")
    parser.add_argument('-df', '--debugging_folder', default='logs/', type=str, dest="debugging_folder")
    parser.add_argument('-rs', '--random_start', default=True, type=bool_arg, dest="random_start")
    parser.add_argument('--scale', default=1000., type=float)
    parser.add_argument('--height', default=84, type=int)
    parser.add_argument('--filters', default=32, type=int)
    parser.add_argument('--rnn-length', default=5, type=int)
    parser.add_argument('--static-size', default=2, type=int)
    parser.add_argument('--temporal-size', default=2, type=int)
    parser.add_argument('--static-hidden-size', default=32, type=int)
    parser.add_argument('--temporal-hidden-size', default=32, type=int)
    # additions
    parser.add_argument('--n_locusts', default=None, type=int, help="SwarmEnv.N_LOCUSTS (default 80)")
    parser.add_argument('--max_updates', default=None, type=int)
    parser.add_argument('--no_cuda_graph', action='store_true')
    parser.add_argument('--reward_indexing', default='reference', choices=['reference', 'per_agent'])
    parser.add_argument('--mask_terminals', action='store_true')
    parser.add_argument('--obs', default='auto', choices=['auto', 'compact', 'expanded'],
                        help="observation the net consumes: compact = (grid, positions) with conv1 factorised")
    return parser


def main(args):
    import torch
    import torch.distributed as dist
    import golds_rl_gym_b200 as pkg
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.device.startswith('/gpu:') and world == 1:
        local = int(args.device.split(':')[1])
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    network_creator, env_creator = get_network_and_environment_creator(args)
    paac = pkg.submodule("agents.paac.paac")
    learner = paac.GridPAACLearner(network_creator, env_creator, args, reward_indexing=args.reward_indexing,
                                   mask_terminals=args.mask_terminals, use_cuda_graph=not args.no_cuda_graph,
                                   compact_obs={"auto": "auto", "compact": True, "expanded": False}[args.obs])

    def on_signal(signum, frame):            # train_paac_conv.py:47-58
        learner.cleanup()
        sys.exit(0)
    signal.signal(signal.SIGINT, on_signal)
    signal.signal(signal.SIGTERM, on_signal)
    fps = learner.train(max_updates=args.max_updates)
    if int(os.environ.get("RANK", "0")) == 0:
        logging.info("done: %d global steps, %.1f frames/s", learner.global_step, fps)
    learner.cleanup()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main(get_arg_parser().parse_args())
