"""CPU model of K3's schedule (nb_resolve.cu): the reference handles its event queue serially in
reverse arrival order (PushFront + Front→Next, cmd/body/body_collection.go:82-104,212-233), i.e. in
descending key (i << 32 | j); K3 runs rounds in which every event that holds the largest pending
key of BOTH of its bodies is resolved in parallel.  This checks the claim the kernel rests on, for
arbitrary event lists and a deliberately non-commutative update: the rounds touch disjoint bodies,
terminate, and leave exactly the state of the serial order — whatever order K1 appended the events in.
"""
import random

from hypothesis import given, settings
from hypothesis import strategies as st


def key(i, j):
    return (i << 32) | j


def apply_event(state, i, j, k):
    # non-commutative, order-sensitive toy "collision": any deviation from the serial order shows
    a, b = state[i], state[j]
    state[i] = (a * 31 + b * 17 + k) % 1_000_003
    state[j] = (b * 29 + a * 13 + 2 * k + 1) % 1_000_003


def serial(n, events):
    state = list(range(1, n + 1))
    for i, j in sorted(set(events), key=lambda e: key(*e), reverse=True):
        apply_event(state, i, j, key(i, j) % 9973)
    return state


def wavefront(n, events, shuffle_seed):
    state = list(range(1, n + 1))
    pending = list(set(events))
    random.Random(shuffle_seed).shuffle(pending)        # K1 appends with atomicAdd: any order
    rounds = 0
    while pending:
        head = {}
        for i, j in pending:                              # largest pending key per body
            k = key(i, j)
            head[i] = max(head.get(i, -1), k)
            head[j] = max(head.get(j, -1), k)
        ready = [(i, j) for i, j in pending if head[i] == key(i, j) and head[j] == key(i, j)]
        assert ready, "the event with the globally largest key is always ready"
        touched = [b for e in ready for b in e]
        assert len(touched) == len(set(touched)), "ready events touch disjoint bodies"
        for i, j in ready:                                # any order inside a round
            apply_event(state, i, j, key(i, j) % 9973)
        done = set(ready)
        pending = [e for e in pending if e not in done]
        rounds += 1
    return state, rounds


pairs = st.integers(2, 40).flatmap(lambda n: st.tuples(
    st.just(n),
    st.lists(st.tuples(st.integers(0, n - 1), st.integers(0, n - 1)).filter(lambda e: e[0] != e[1]), max_size=120),
    st.integers(0, 1 << 30)))


@settings(max_examples=300, deadline=None, derandomize=True)
@given(pairs)
def test_wavefront_rounds_equal_the_serial_reverse_arrival_order(case):
    n, events, seed = case
    state, rounds = wavefront(n, events, seed)
    assert state == serial(n, events)
    assert rounds <= max(len(set(events)), 0)


def test_symmetric_pairs_resolve_in_two_rounds_and_chains_in_their_length():
    # C4-like: isolated overlapping pairs emit (i,j) and (j,i) -> two rounds (bench: resolve_rounds = 2)
    ev = [(2 * k, 2 * k + 1) for k in range(50)] + [(2 * k + 1, 2 * k) for k in range(50)]
    state, rounds = wavefront(100, ev, 1)
    assert rounds == 2 and state == serial(100, ev)
    # a chain 0-1, 1-2, ..., 29-30 in one direction is one dependency chain: 30 rounds
    chain = [(k, k + 1) for k in range(30)]
    state, rounds = wavefront(31, chain, 2)
    assert rounds == 30 and state == serial(31, chain)


def short_list_schedule(n, events, shuffle_seed):
    """The shared-memory schedule K3 uses for lists of up to 64 events (nb_resolve.cu, `FAST_MAX`): no per-body heads;
    an event is ready when no pending event with a LARGER key shares one of its bodies.  Returns the state and the
    list of rounds (each a set of events)."""
    state = list(range(1, n + 1))
    pending = list(set(events))
    random.Random(shuffle_seed).shuffle(pending)
    rounds = []
    while pending:
        ready = [e for e in pending
                 if not any(key(*o) > key(*e) and (o[0] in e or o[1] in e) for o in pending)]
        assert ready
        for i, j in ready:
            apply_event(state, i, j, key(i, j) % 9973)
        rounds.append(set(ready))
        pending = [e for e in pending if e not in rounds[-1]]
    return state, rounds


def wavefront_rounds(n, events):
    pending, rounds = set(events), []
    while pending:
        head = {}
        for i, j in pending:
            head[i] = max(head.get(i, -1), key(i, j))
            head[j] = max(head.get(j, -1), key(i, j))
        ready = {(i, j) for i, j in pending if head[i] == key(i, j) and head[j] == key(i, j)}
        rounds.append(ready)
        pending -= ready
    return rounds


@settings(max_examples=300, deadline=None, derandomize=True)
@given(pairs)
def test_short_list_rule_makes_the_same_rounds_as_the_heads(case):
    """Same events in the same rounds (so `resolve_rounds` and every result bit agree), same final state as the
    serial order — the GPU test compares the two kernel paths, this one the two rules."""
    n, events, seed = case
    events = events[:64]
    state, rounds = short_list_schedule(n, events, seed)
    assert rounds == wavefront_rounds(n, events)
    assert state == serial(n, events)
