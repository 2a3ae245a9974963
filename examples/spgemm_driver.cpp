// spgemm_driver.cpp -- the reference's benchmark driver (SpGEMM_cuda/main.cu) rebuilt on
// the drop-in class, with the same command line:
//
//     ./spgemm -cuda -spgemm 0            hand-built 4x6 * 6x4 test   (main.cu:149-246)
//     ./spgemm -cuda -spgemm 1|2|3|4      Poisson 5pt 256^2, 9pt 256^2, 7pt 51^3, 27pt 51^3 (main.cu:30-53)
//     ./spgemm -cuda -spgemm A.mtx [B.mtx]
//
// CUSP (gallery, MatrixMarket reader, cusp::multiply check) is replaced by the small
// generators / reader below and by a host Gustavson check with the reference's comparison
// (nnzC, rowptr, columns exact, values within 10 %: ref_spgemm.h:79-126).
// build: python -m benchmark_spgemm_using_csr_b200.build --driver   (g++ -I include, links libbhsparse_b200.so)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "bhsparse.h"

using namespace std;

struct Csr {
    int rows = 0, cols = 0;
    vector<index_type> rowptr, col;
    vector<value_type> val;
};

static Csr stencil(int nx, int ny, int nz, bool box)
{
    Csr A;
    const int N = nx * ny * nz;
    A.rows = A.cols = N;
    A.rowptr.assign(1, 0);
    for (int z = 0; z < nz; z++)
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) {
                for (int dz = -1; dz <= 1; dz++)
                    for (int dy = -1; dy <= 1; dy++)
                        for (int dx = -1; dx <= 1; dx++) {
                            if (!box && (abs(dx) + abs(dy) + abs(dz) > 1)) continue;
                            const int X = x + dx, Y = y + dy, Z = z + dz;
                            if (X < 0 || X >= nx || Y < 0 || Y >= ny || Z < 0 || Z >= nz) continue;
                            A.col.push_back((Z * ny + Y) * nx + X);
                        }
                A.rowptr.push_back((int)A.col.size());
            }
    A.val.assign(A.col.size(), 1.0);
    return A;
}

// MatrixMarket coordinate reader (main.cu:56-64 uses cusp::io::read_matrix_market_file): real /
// integer / pattern fields, general / symmetric / skew-symmetric; duplicates summed, columns sorted.
// Returns false (with a message) on anything else or on a malformed file.
static bool read_mtx(const string &path, Csr &A)
{
    ifstream f(path);
    if (!f) {
        cout << "cannot open " << path << endl;
        return false;
    }
    string line, banner, object, format, field, symmetry;
    getline(f, line);
    istringstream hdr(line);
    hdr >> banner >> object >> format >> field >> symmetry;
    auto lower = [](string &t) { for (auto &ch : t) ch = (char)tolower((unsigned char)ch); };
    lower(object), lower(format), lower(field), lower(symmetry);
    if (banner != "%%MatrixMarket" || object != "matrix" || format != "coordinate") {
        cout << path << ": only MatrixMarket coordinate matrices are supported" << endl;
        return false;
    }
    const bool pattern = field == "pattern";
    if (!pattern && field != "real" && field != "integer" && field != "double") {
        cout << path << ": unsupported field '" << field << "'" << endl;
        return false;
    }
    double mirror = 0.0;   // value factor of the mirrored entry; 0 = general
    if (symmetry == "symmetric") mirror = 1.0;
    else if (symmetry == "skew-symmetric") mirror = -1.0;
    else if (symmetry != "general") {
        cout << path << ": unsupported symmetry '" << symmetry << "'" << endl;
        return false;
    }
    while (getline(f, line) && (line.empty() || line[0] == '%')) {}
    long rows = 0, cols = 0, nnz = -1;
    istringstream(line) >> rows >> cols >> nnz;
    if (rows < 0 || cols < 0 || nnz < 0 || rows > 0x7fffffffL || cols > 0x7fffffffL) {
        cout << path << ": bad size line" << endl;
        return false;
    }
    vector<map<int, double>> R((size_t)rows);
    for (long e = 0; e < nnz; e++) {
        long i = 0, j = 0;
        double v = 1.0;
        if (!(f >> i >> j) || (!pattern && !(f >> v))) {
            cout << path << ": truncated entry list" << endl;
            return false;
        }
        if (i < 1 || i > rows || j < 1 || j > cols) {
            cout << path << ": entry (" << i << ", " << j << ") outside the matrix" << endl;
            return false;
        }
        R[i - 1][(int)(j - 1)] += v;
        if (mirror != 0.0 && i != j) {
            if (j > rows || i > cols) {
                cout << path << ": symmetric entry outside the matrix" << endl;
                return false;
            }
            R[j - 1][(int)(i - 1)] += mirror * v;
        }
    }
    A.rows = (int)rows;
    A.cols = (int)cols;
    A.rowptr.assign(1, 0);
    for (auto &r : R) {   // std::map: columns ascending == csr_sort_indices (ref_spgemm.h:37-62)
        for (auto &kv : r) {
            A.col.push_back(kv.first);
            A.val.push_back((value_type)kv.second);
        }
        A.rowptr.push_back((int)A.col.size());
    }
    return true;
}

// host check with the reference's comparison semantics (ref_spgemm.h:79-126)
static void compData(const Csr &A, const Csr &B, int nnzC, const index_type *rpC, const index_type *cC, const value_type *vC)
{
    cout << endl << "Checking correctness ..." << endl;
    vector<index_type> rp(A.rows + 1, 0), cc;
    vector<value_type> vv;
    vector<int> mark(B.cols, -1);
    vector<value_type> acc(B.cols);
    vector<int> touched;
    for (int i = 0; i < A.rows; i++) {
        touched.clear();
        for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; p++)
            for (int q = B.rowptr[A.col[p]]; q < B.rowptr[A.col[p] + 1]; q++) {
                const int j = B.col[q];
                if (mark[j] != i) {
                    mark[j] = i;
                    acc[j] = 0;
                    touched.push_back(j);
                }
                acc[j] += A.val[p] * B.val[q];
            }
        sort(touched.begin(), touched.end());
        for (int j : touched) {
            cc.push_back(j);
            vv.push_back(acc[j]);
        }
        rp[i + 1] = (int)cc.size();
    }
    if ((int)cc.size() == nnzC)
        cout << "nnzC = " << nnzC << ". PASS!" << endl;
    else {
        cout << "nnzC = " << nnzC << ", reference nnzC = " << cc.size() << ". NO PASS!" << endl;
        return;
    }
    int err = 0;
    for (int i = 0; i <= A.rows; i++) err += rp[i] != rpC[i];
    cout << (err ? "RowPtrC NO PASS!" : "RowPtrC PASS!") << endl;
    if (err) return;
    err = 0;
    for (int j = 0; j < nnzC; j++)
        if (cc[j] != cC[j] || fabs((double)vv[j] - (double)vC[j]) > fabs(0.1 * (double)vv[j])) err++;
    if (!err)
        cout << "ColIndC/csrValC PASS!" << endl;
    else
        cout << "ColIndC/csrValC NO PASS! #err = " << err << endl;
}

static int run(Csr &A, Csr &B, bool *platforms, int warmups)
{
    cout << " A: ( " << A.rows << " by " << A.cols << ", nnz = " << A.col.size() << " ) " << endl;
    cout << " B: ( " << B.rows << " by " << B.cols << ", nnz = " << B.col.size() << " ) " << endl;
    vector<index_type> rowptrC(A.rows + 1);
    int err = 0;
    std::unique_ptr<bhsparse> bh_sparse(new bhsparse());   // call sequence of main.cu:104-135 (released on every return)
    err = bh_sparse->initPlatform(platforms);
    if (err != BHSPARSE_SUCCESS) return err;
    err = bh_sparse->initData(A.rows, A.cols, B.cols, (int)A.col.size(), A.val.data(), A.rowptr.data(), A.col.data(),
                              (int)B.col.size(), B.val.data(), B.rowptr.data(), B.col.data(), rowptrC.data());
    if (err != BHSPARSE_SUCCESS) return err;
    for (int i = 0; i < warmups; i++) {
        err = bh_sparse->warmup();
        if (err != BHSPARSE_SUCCESS) return err;
    }
    err = bh_sparse->spgemm();
    if (err != BHSPARSE_SUCCESS) return err;
    int nnzC = bh_sparse->get_nnzC();
    vector<index_type> colC(max(nnzC, 1));
    vector<value_type> valC(max(nnzC, 1));
    err = bh_sparse->get_C(colC.data(), valC.data());
    if (err != BHSPARSE_SUCCESS) return err;
    err = bh_sparse->free_mem();
    if (err != BHSPARSE_SUCCESS) return err;
    err = bh_sparse->freePlatform();
    if (err != BHSPARSE_SUCCESS) return err;
    compData(A, B, nnzC, rowptrC.data(), colC.data(), valC.data());
    return BHSPARSE_SUCCESS;
}

static int test_small_spgemm(bool *platforms)   // main.cu:149-246
{
    Csr A, B;
    A.rows = 4;
    A.cols = 6;
    A.rowptr = {0, 1, 4, 5, 6};
    A.col = {0, 1, 2, 3, 3, 1};
    for (int i = 0; i < 6; i++) A.val.push_back((value_type)((i + 1) * 10));
    B.rows = 6;
    B.cols = 4;
    B.rowptr = {0, 1, 3, 5, 5, 5, 7};
    B.col = {0, 1, 3, 0, 1, 1, 3};
    for (int i = 0; i < 7; i++) B.val.push_back((value_type)(i + 1));
    return run(A, B, platforms, 0);
}

static int benchmark_spgemm(const char *d1, const char *d2, bool *platforms)   // main.cu:23-147
{
    Csr A, B;
    if (!strcmp(d1, "1")) { A = stencil(256, 256, 1, false); B = A; cout << "2D FD, 5-point. "; }
    else if (!strcmp(d1, "2")) { A = stencil(256, 256, 1, true); B = A; cout << "2D FE, 9-point. "; }
    else if (!strcmp(d1, "3")) { A = stencil(51, 51, 51, false); B = A; cout << "3D FD, 7-point. "; }
    else if (!strcmp(d1, "4")) { A = stencil(51, 51, 51, true); B = A; cout << "3D FE, 27-point. "; }
    else {
        cout << " A: " << d1 << endl;
        if (!read_mtx(d1, A)) { cout << "cannot read " << d1 << endl; return -1; }
        cout << " B: " << d2 << endl;
        if (!read_mtx(d2, B)) { cout << "cannot read " << d2 << endl; return -1; }
    }
    if (A.cols != B.rows) { cout << "inner dimensions differ" << endl; return -1; }
    srand(1);   // main.cu:79 seeds with time(NULL); fixed here so runs repeat
    for (auto &v : A.val) v = (value_type)((rand() % 9) + 1);
    for (auto &v : B.val) v = (value_type)((rand() % 9) + 1);
    return run(A, B, platforms, 3);
}

int main(int argc, char **argv)   // main.cu:248-314
{
    bool platforms[NUM_PLATFORMS];
    memset(platforms, 0, sizeof(platforms));
    int argi = 1;
    const char *d1 = nullptr, *d2 = nullptr;
    if (argc > argi) {
        if (!strcmp(argv[argi], "-cuda")) platforms[BHSPARSE_CUDA] = true;
        argi++;
    }
    if (argc > argi && !strcmp(argv[argi], "-spgemm")) {
        argi++;
        if (argc > argi) d1 = argv[argi++];
        d2 = (argc > argi) ? argv[argi++] : d1;
    }
    if (!d1) {
        cout << "usage: spgemm -cuda -spgemm <0|1|2|3|4|A.mtx> [B.mtx]" << endl;
        return 1;
    }
    cout << "------------------------" << endl;
    int err = !strcmp(d1, "0") ? test_small_spgemm(platforms) : benchmark_spgemm(d1, d2, platforms);
    if (err != BHSPARSE_SUCCESS) cout << "Found an err, code = " << err << endl;
    cout << "------------------------" << endl;
    return err != BHSPARSE_SUCCESS;
}
