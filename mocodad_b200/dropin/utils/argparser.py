"""``from utils.argparser import init_args`` (eval_MoCoDAD.py:8): the path derivation of utils/argparser.py:4-43 -- ground-truth
directory, pose directories per split, experiment directory ``{exp_dir}/{dataset_choice}/{dir_name}`` -- on the YAML namespace."""
import os


def create_experiment_dirs(args) -> str:
    """utils/argparser.py:31-43"""
    checkpoints_dir = os.path.join(args.exp_dir, args.dataset_choice, args.dir_name)
    if args.create_experiment_dir:
        try:
            os.makedirs(checkpoints_dir, exist_ok=True)
            print("Experiment directories created in {}".format(checkpoints_dir))
        except Exception as err:
            print("Experiment directories creation Failed, error {}".format(err))
            exit(-1)
    return checkpoints_dir


def init_args(args):
    """utils/argparser.py:4-28"""
    if args.debug:
        args.ae_epochs = 10
    args.gt_path = args.test_path
    pose = os.path.join(args.data_dir, 'pose')
    if args.dataset_choice in ['STC', 'HR-STC', 'HR-Avenue', 'UBnormal']:
        args.pose_path = {'train': os.path.join(pose, 'training/tracked_person/'),
                          'test': os.path.join(pose, 'testing/tracked_person/'),
                          'validation': os.path.join(pose, 'validating/tracked_person/')}
        if args.validation:
            args.gt_path = os.path.join(args.data_dir, 'validating', 'test_frame_mask')
    elif args.dataset_choice == 'Avenue':
        print('Not usable yet.')
        exit(-1)
    args.ckpt_dir = create_experiment_dirs(args)
    return args
