"""Host-side model of the tile schedule of csrc/dct_tile.cuh (static ranges + tail pool, with the DCT_POOL_SPLIT
sub-tiles): replays draw() for every CTA against one shared counter in an adversarial interleaving and checks that every
pixel of every image is handed out exactly once.  A design check of the index arithmetic -- the kernel itself is checked by
the GPU parity suite against each variant library (tools/ab_pool_split.sh)."""
import itertools
import random


def schedule(num_images, HW, TP, grid, pool_div, S, L, seed):
    tpi = (HW + TP - 1) // TP
    num_tiles = tpi * num_images
    grid = min(grid, num_tiles)
    per, extra = divmod(num_tiles, grid)
    pool_q = per // pool_div
    TPS = TP // S
    counter = [0]

    def span(t, part):
        b = t // tpi
        off = (t - b * tpi) * TP
        ln = min(HW - off, TP)
        if part >= 0:
            o = part * TPS
            off += o
            ln = max(0, min(ln - o, TPS))
        return b, off, ln

    def draw(cta, st):
        my_begin = cta * per + min(cta, extra)
        my_n = per + (1 if cta < extra else 0)
        my_static = my_n - pool_q
        if st["draws"] >= my_static and pool_q != 0 and S > 1:
            lv = min(pool_q, L)
            whole = pool_q - lv
            n_whole = whole * grid
            while True:
                got = counter[0]; counter[0] += 1
                part = -1
                if got < n_whole:
                    j, k = got % grid, got // grid
                else:
                    r = got - n_whole
                    w = r % (grid * S)
                    k = whole + r // (grid * S); j = w % grid; part = w // grid
                if k >= pool_q:
                    t = num_tiles; break
                t = j * per + min(j, extra) + per + (1 if j < extra else 0) - pool_q + k
                if span(t, part)[2] > 0:
                    break
            st["draws"] += 1
            return t, part
        if st["draws"] < my_static:
            t = my_begin + st["draws"]
        elif pool_q == 0:
            t = num_tiles
        else:
            got = counter[0]; counter[0] += 1
            j, k = got % grid, got // grid
            t = j * per + min(j, extra) + per + (1 if j < extra else 0) - pool_q + k if k < pool_q else num_tiles
        st["draws"] += 1
        return t, -1

    rng = random.Random(seed)
    cover = [[0] * HW for _ in range(num_images)]
    live = {c: {"draws": 0} for c in range(grid)}
    while live:
        c = rng.choice(list(live))
        t, part = draw(c, live[c])
        if t >= num_tiles:
            del live[c]
            continue
        b, off, ln = span(t, part)
        assert ln > 0
        for p in range(off, off + ln):
            cover[b][p] += 1
    assert all(v == 1 for img in cover for v in img), "a pixel was handed out %s times" % \
        sorted({v for img in cover for v in img})


if __name__ == "__main__":
    n = 0
    for (B, HW, TP, grid, pd, S, L) in itertools.product([1, 3], [512, 1000, 4100], [64, 160], [1, 7, 296], [1, 2, 5],
                                                       [1, 2, 4], [1, 2, 9]):
        if TP % S or HW % 4:
            continue
        schedule(B, HW, TP, grid, pd, S, L, seed=n)
        n += 1
    print(f"{n} schedules: every pixel handed out exactly once")
