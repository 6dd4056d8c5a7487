#!/usr/bin/env python
"""Drop-in for the reference's evaluate_retrieval.py CLI (evaluate_retrieval.py:155-208): same flags, same
feature pickles, same metrics table -- the all-pairs distance matrix is computed by the CUDA kernel
(semantic_embeddings_b200.evaluate_retrieval), the hierarchical precision by the reference's own
ClassHierarchy (class_hierarchy.py, imported from --reference_root: it is pure Python/numpy and not part of
the accelerated hot path; SURVEY.md section 8f ranks it as the next row)."""
import argparse
import os
import pickle
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from semantic_embeddings_b200.evaluate_retrieval import pairwise_retrieval  # noqa: E402

METRICS = ['P@1 (WUP)', 'P@10 (WUP)', 'P@50 (WUP)', 'P@100 (WUP)', 'AHP (WUP)', 'P@1 (LCS_HEIGHT)', 'P@10 (LCS_HEIGHT)',
           'P@50 (LCS_HEIGHT)', 'P@100 (LCS_HEIGHT)', 'AHP (LCS_HEIGHT)', 'AP']


def print_performance(perf, metrics=METRICS):
    """evaluate_retrieval.py:76-100 (same table layout)."""
    print()
    max_name_len = max(len(lbl) for lbl in perf.keys())
    print(' | '.join([' ' * max_name_len] + ['{:^6s}'.format(metric) for metric in metrics]))
    print('-' * (max_name_len + sum(3 + max(len(metric), 6) for metric in metrics)))
    for lbl, metric_values in perf.items():
        print('{:{}s} | {}'.format(lbl, max_name_len, ' | '.join(
            '{:>{}.4f}'.format(metric_values[m], max(len(m), 6)) if m in metric_values else ' ' * max(len(m), 6)
            for m in metrics)))
    print()


def write_performance(perf, csv_file, metrics=METRICS):
    """evaluate_retrieval.py:103-113."""
    with open(csv_file, 'w') as f:
        f.write(';'.join([''] + metrics) + '\n')
        for lbl, vals in perf.items():
            f.write(';'.join([lbl] + ['{:.6f}'.format(vals[m]) if m in vals else '' for m in metrics]) + '\n')


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='Evaluates content-based image retrieval performance.',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--dataset', type=str, required=True)
    parser.add_argument('--data_root', type=str, required=True)
    parser.add_argument('--hierarchy', type=str, required=True)
    parser.add_argument('--is_a', action='store_true', default=False)
    parser.add_argument('--str_ids', action='store_true', default=False)
    parser.add_argument('--classes_from', type=str, default=None)
    parser.add_argument('--feat', type=str, action='append', required=True)
    parser.add_argument('--label', type=str, action='append')
    parser.add_argument('--norm', type=str, action='append')
    parser.add_argument('--plot_max', type=int, default=250)
    parser.add_argument('--prec_type', type=str, default='LCS_HEIGHT', choices=['LCS_HEIGHT', 'WUP'])
    parser.add_argument('--clip_ahp', type=int, default=None)
    parser.add_argument('--csv', type=str, default=None)
    parser.add_argument('--reference_root', type=str, default=os.environ.get('SEMANTIC_EMBEDDINGS_REFERENCE', '/root/reference'),
                        help='(new) checkout of cvjena/semantic-embeddings providing class_hierarchy.py')
    args = parser.parse_args()
    sys.path.insert(0, args.reference_root)
    from class_hierarchy import ClassHierarchy
    from learn_image_embeddings import get_data_generator

    id_type = str if args.str_ids else int
    if args.classes_from:
        with open(args.classes_from, 'rb') as pf:
            embed_labels = pickle.load(pf)['ind2label']
    else:
        embed_labels = None
    data_generator = get_data_generator(args.dataset, args.data_root, classes=embed_labels)
    labels_test = data_generator.labels_test
    if embed_labels is not None:
        labels_test = [embed_labels[lbl] for lbl in labels_test]
    labels_test = dict(enumerate(labels_test))
    hierarchy = ClassHierarchy.from_file(args.hierarchy, is_a_relations=args.is_a, id_type=id_type)

    perf = OrderedDict()
    for i, feat_dump in enumerate(args.feat):
        label = args.label[i] if args.label and i < len(args.label) else os.path.splitext(os.path.basename(feat_dump))[0]
        normalize = bool(args.norm and i < len(args.norm) and args.norm[i].lower() in ('yes', 'y', 'true', '1'))
        retrieved = pairwise_retrieval(feat_dump, normalize)
        perf[label], _ = hierarchy.hierarchical_precision(
            retrieved, labels_test, ks=[1, 10, 50, 100], compute_ahp=args.clip_ahp if args.clip_ahp else True,
            compute_ap=True, all_ids=list(labels_test.keys()))
    metrics = METRICS if not args.clip_ahp else [m.replace('AHP', 'AHP@{}'.format(args.clip_ahp)) for m in METRICS]
    print_performance(perf, metrics)
    if args.csv:
        write_performance(perf, args.csv, metrics)
