"""Host-side retrieval logic that needs no GPU: the candidate merge used for k > 32."""
import torch

from nans_clip_b200.retrieval import _merge_candidates


def test_merge_candidates_is_the_reference_order():
    """(score descending, gallery position ascending) = Python's stable sorted(..., reverse=True) over
    tuples listed in gallery order (make_topk_predictions.py:84); index -1 entries are padding."""
    g = torch.Generator().manual_seed(0)
    Q, C, k = 7, 40, 12
    idx = torch.stack([torch.randperm(1000, generator=g)[:C] for _ in range(Q)])
    sc = torch.randint(0, 6, (Q, C), generator=g).float() / 4        # many exact ties
    idx[0, 5:9] = -1                                                   # padding entries anywhere in the list
    s, i = _merge_candidates(sc.clone(), idx.clone(), k)
    for q in range(Q):
        tup = sorted([(int(idx[q, c]), float(sc[q, c])) for c in range(C) if int(idx[q, c]) >= 0])   # gallery order
        want = sorted(tup, key=lambda x: x[1], reverse=True)[:k]
        assert i[q].tolist() == [t[0] for t in want] and s[q].tolist() == [t[1] for t in want]
    # fewer candidates than k: padded with (-inf, -1)
    s, i = _merge_candidates(torch.tensor([[0.5, 0.25]]), torch.tensor([[3, -1]]), 4)
    assert i.tolist() == [[3, -1, -1, -1]] and s[0, 0] == 0.5 and torch.isinf(s[0, 1:]).all()
