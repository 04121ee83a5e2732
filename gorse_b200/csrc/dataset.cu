// dataset.cu -- the NCF-format loader behind `gorse-bench cf`: dataset.LoadDataFromBuiltIn's two parsers,
// dataset/dataset.go:398-490, producing the CSR arrays gorse_b200_cf_create / gorse_b200_eval_create take.  Host code.
//   train file (loadTrain, :420-453): one "user<TAB>item[<TAB>...]" per line, integer ids; users and items are created for
//       every id up to the largest seen (:432-444), feedback is appended in file order (duplicates kept, dataset.go:231-240)
//   test file (loadTest, :455-490): "(user,item)<TAB>neg<TAB>neg..." -- the held-out positive of the user and its sampled
//       negatives; a negative that names an unseen item adds it to the item dictionary (:485, itemDict.Add)
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "common.cuh"

struct gorse_b200_ncf {
    int32_t n_users = 0, n_items = 0;
    std::vector<std::vector<int32_t>> train, test, neg;   // per user
};

namespace {

bool parse_i32(const char *b, const char *e, int32_t *out)
{
    if (b == e) return false;
    char *end = nullptr;
    errno = 0;
    const std::string s(b, e);
    long v = strtol(s.c_str(), &end, 10);
    if (errno || *end != '\0' || v < 0 || v > 0x7fffffffL) return false;
    *out = (int32_t)v;
    return true;
}

// getline-based reader without a line-length limit (bufio.Scanner would fail beyond 64 KB; real NCF test lines are ~1 KB)
bool read_line(FILE *f, std::string &line)
{
    line.clear();
    int ch;
    while ((ch = fgetc(f)) != EOF) {
        if (ch == '\n') return true;
        line.push_back((char)ch);
    }
    return !line.empty();
}

void flatten(const std::vector<std::vector<int32_t>> &rows, int32_t n_rows, int64_t *off, int32_t *idx)
{
    int64_t p = 0;
    for (int32_t r = 0; r < n_rows; r++) {
        if (off) off[r] = p;
        if (r < (int32_t)rows.size()) {
            if (idx) for (int32_t v : rows[(size_t)r]) idx[p++] = v;
            else p += (int64_t)rows[(size_t)r].size();
        }
    }
    if (off) off[n_rows] = p;
}

int64_t total(const std::vector<std::vector<int32_t>> &rows)
{
    int64_t n = 0;
    for (auto &r : rows) n += (int64_t)r.size();
    return n;
}

}  // namespace

extern "C" {

int32_t gorse_b200_ncf_free(gorse_b200_ncf *d)
{
    delete d;
    return GORSE_B200_OK;
}

int32_t gorse_b200_ncf_load(const char *train_path, const char *test_path, gorse_b200_ncf **out)
{
    GB_CHECK_ARG(train_path != nullptr && out != nullptr, "NULL argument");
    *out = nullptr;
    FILE *f = fopen(train_path, "r");
    if (!f) { gb::set_error("open %s: %s", train_path, strerror(errno)); return GORSE_B200_ERR_ARG; }
    gorse_b200_ncf *d = new gorse_b200_ncf();
    std::string line;
    int64_t ln = 0;
    auto fail = [&](const char *path) {
        gb::set_error("%s:%lld: wrong format: %s", path, (long long)ln, line.substr(0, 80).c_str());
        fclose(f);
        delete d;
        return GORSE_B200_ERR_ARG;
    };
    while (read_line(f, line)) {
        ln++;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        const char *b = line.data(), *e = b + line.size();
        const char *t1 = (const char *)memchr(b, '\t', (size_t)(e - b));
        if (!t1) return fail(train_path);
        const char *t2 = (const char *)memchr(t1 + 1, '\t', (size_t)(e - t1 - 1));
        int32_t u, i;
        if (!parse_i32(b, t1, &u) || !parse_i32(t1 + 1, t2 ? t2 : e, &i)) return fail(train_path);
        d->n_users = std::max(d->n_users, u + 1);       // :432-436
        d->n_items = std::max(d->n_items, i + 1);       // :441-444
        if ((int32_t)d->train.size() < d->n_users) d->train.resize((size_t)d->n_users);
        d->train[(size_t)u].push_back(i);               // :446
    }
    fclose(f);
    d->test.resize((size_t)d->n_users);
    d->neg.resize((size_t)d->n_users);
    if (test_path) {
        f = fopen(test_path, "r");
        if (!f) { gb::set_error("open %s: %s", test_path, strerror(errno)); delete d; return GORSE_B200_ERR_ARG; }
        ln = 0;
        while (read_line(f, line)) {
            ln++;
            if (!line.empty() && line.back() == '\r') line.pop_back();
            const char *b = line.data(), *e = b + line.size();
            const char *t1 = (const char *)memchr(b, '\t', (size_t)(e - b));
            const char *pe = t1 ? t1 : e;
            if (pe - b < 5 || b[0] != '(' || pe[-1] != ')') return fail(test_path);      // :468-470
            const char *comma = (const char *)memchr(b, ',', (size_t)(pe - b));
            int32_t u, i;
            if (!comma || !parse_i32(b + 1, comma, &u) || !parse_i32(comma + 1, pe - 1, &i)) return fail(test_path);
            if (u >= d->n_users) {      // the reference's fixed-size slices would panic here; grow instead
                d->n_users = u + 1;
                d->train.resize((size_t)d->n_users); d->test.resize((size_t)d->n_users); d->neg.resize((size_t)d->n_users);
            }
            d->n_items = std::max(d->n_items, i + 1);
            d->test[(size_t)u].push_back(i);            // :474
            std::vector<int32_t> &ng = d->neg[(size_t)u];
            ng.clear();                                 // :480: the last line of a user wins
            const char *p = t1;
            while (p && p < e) {
                const char *q = (const char *)memchr(p + 1, '\t', (size_t)(e - p - 1));
                int32_t v;
                if (!parse_i32(p + 1, q ? q : e, &v)) return fail(test_path);
                d->n_items = std::max(d->n_items, v + 1);   // itemDict.Add of an unseen id (:485)
                ng.push_back(v);
                p = q;
            }
        }
        fclose(f);
    }
    *out = d;
    return GORSE_B200_OK;
}

int32_t gorse_b200_ncf_shape(const gorse_b200_ncf *d, int32_t *n_users, int32_t *n_items, int64_t *n_train, int64_t *n_test, int64_t *n_neg)
{
    GB_CHECK_ARG(d != nullptr, "dataset is NULL");
    if (n_users) *n_users = d->n_users;
    if (n_items) *n_items = d->n_items;
    if (n_train) *n_train = total(d->train);
    if (n_test) *n_test = total(d->test);
    if (n_neg) *n_neg = total(d->neg);
    return GORSE_B200_OK;
}

int32_t gorse_b200_ncf_get(const gorse_b200_ncf *d, int64_t *train_off, int32_t *train_items, int64_t *test_off, int32_t *test_items,
                           int64_t *neg_off, int32_t *neg_items)
{
    GB_CHECK_ARG(d != nullptr, "dataset is NULL");
    flatten(d->train, d->n_users, train_off, train_items);
    flatten(d->test, d->n_users, test_off, test_items);
    flatten(d->neg, d->n_users, neg_off, neg_items);
    return GORSE_B200_OK;
}

}  // extern "C"
