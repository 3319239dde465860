// TEST INFRASTRUCTURE: the device routine of pcrcg_b200/csrc/label_vote.h compiled for the host (g++), so that its logic can be
// compared with the live std::unordered_map of oracle/_ref without a GPU (tests/test_label_vote_host.py).
#include "../../pcrcg_b200/csrc/label_vote.h"

extern "C" int host_label_vote(const int* labels, long n, int* overflow)
{
    pcrcg::LabelVote v;
    v.reset();
    for (long i = 0; i < n; i++) v.add(labels[i]);
    *overflow = v.overflow ? 1 : 0;
    return v.pick();
}
