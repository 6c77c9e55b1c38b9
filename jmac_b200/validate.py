"""Drop-in for ``CompletionEvaluator`` of the reference's ``src/validate.py`` (filtered link
prediction: Hits@1, Hits@10, MRR).

Same constructor and ``test`` signature / return value.  The reference negates the [B,N] distance
matrix, overwrites the true tails with -1e6 in a per-row Python loop with ``.item()`` syncs, sorts
every row fully and looks the target up (src/validate.py:50-70).  Here the rank of the target is
obtained by counting candidates strictly closer than the target (``jmac_l1_rank``), with the
filter passed as a CSR list -- no [B,N] matrix, no sort, no per-row host sync.  Ranks are identical
wherever the target's score differs from every competitor's (ties: see DESIGN.md).
"""
from __future__ import division

import numpy as np
import torch

from . import functional as JF


class CompletionEvaluator:
    def __init__(self, target_kg, model, device, data_dir):
        self.target_kg = target_kg
        self.device = device
        self.model = model
        self.data_dir = data_dir
        self.edge_index = torch.LongTensor(target_kg.edge_index).to(device)
        self.edge_type = torch.LongTensor(target_kg.edge_type).to(device)
        self.ent_bases = [target_kg.entity_id_base, target_kg.upper_entity_base]
        self.rel_bases = [target_kg.relation_id_base, target_kg.upper_relation_base]

    def _encode(self):
        _, ents, rels = self.model.forward_base(self.edge_index, self.edge_type, self.ent_bases, self.rel_bases)
        L = self.model.args.num_gcn_layer
        return [e.detach() for e in ents[:L]], [r.detach() for r in rels[:L]]

    def ranks(self, args, is_val=True, filterr=False):
        """1-based ranks of every val/test triple (numpy int64), batch by batch as the reference."""
        er_vocab = self.target_kg.true_tail
        kg = self.target_kg
        if is_val:
            gen = kg.generate_batch_data(kg.h_val, kg.r_val, kg.t_val, batch_size=args.batch_size, shuffle=False)
        else:
            gen = kg.generate_batch_data(kg.h_test, kg.r_test, kg.t_test, batch_size=args.batch_size, shuffle=False)
        cached = None
        out = []
        for batch in gen:
            # the reference re-runs the full-graph encoder for every batch (src/validate.py:50);
            # in eval mode the result is identical, so encode once per test() call
            if cached is None or self.model.training:
                cached = self._encode()
            ents, rels = cached
            hb, rb, tb = batch[:, 0].reshape(-1), batch[:, 1].reshape(-1), batch[:, 2].reshape(-1)
            fptr = fidx = None
            if filterr:
                hl, rl = hb.tolist(), rb.tolist()
                lists = [np.asarray(er_vocab[(hl[j], rl[j])], dtype=np.int64).reshape(-1) for j in range(len(hl))]
                fptr = np.zeros(len(lists) + 1, dtype=np.int64)
                np.cumsum([len(x) for x in lists], out=fptr[1:])
                fidx = np.concatenate(lists) if lists else np.zeros(0, dtype=np.int64)
                if fidx.size == 0:
                    fidx = np.zeros(1, dtype=np.int64)
            rank = JF.l1_rank(ents, rels, hb, rb, tb, fptr, fidx)
            out.append(rank)
        if not out:
            return np.zeros(0, dtype=np.int64)
        return torch.cat(out).cpu().numpy().astype(np.int64)

    def test(self, args, is_val=True, edge_index_valid=None, edge_type_valid=None, filterr=False, logger=None):
        if edge_index_valid is not None:
            self.edge_index = edge_index_valid
            self.edge_type = edge_type_valid
        ranks = self.ranks(args, is_val=is_val, filterr=filterr)
        hits1 = np.mean(ranks <= 1)
        hits10 = np.mean(ranks <= 10)
        mrr = np.mean(1. / ranks)
        if logger is not None:
            logger.info('Hits @1: {0}'.format(hits1))
            logger.info('Hits @10: {0}'.format(hits10))
            logger.info('Mean rank: {0}'.format(np.mean(ranks)))
            logger.info('Mean reciprocal rank: {0}'.format(mrr))
        return hits1, hits10, mrr
