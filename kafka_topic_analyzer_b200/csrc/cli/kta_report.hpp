// kta_report.hpp — the reference's report (src/main.rs:123-179) over the C ABI getters: header block,
// then the 15-column table in prettytable-rs' default format (borders + a separator after every row,
// cells left-aligned with one space of padding), `{:.4}` dirty ratio, chrono's `Display` for DateTime<Utc>.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace kta_report {

struct PartitionRow {  // one table row, src/main.rs:153-171
    int32_t partition;
    int64_t start_offset, end_offset;
    uint64_t total, alive, tombstones;
    float dirty_ratio;
    uint64_t key_null, key_non_null, key_size_sum, value_size_sum, key_size_avg, value_size_avg, message_size_avg;
};

struct Summary {  // src/main.rs:125-143
    std::string topic;
    uint64_t duration_secs, overall_count;
    int64_t earliest_s;
    int32_t earliest_ns;
    int64_t latest_s;
    uint64_t largest_message, smallest_message, overall_size;
    bool has_alive_keys;
    uint64_t alive_keys;
};

// chrono 0.4 `impl Display for DateTime<Utc>`: "YYYY-MM-DD HH:MM:SS[.fff[fff[fff]]] UTC"
inline std::string format_utc(int64_t secs, int32_t nanos) {
    int64_t days = secs / 86400, rem = secs % 86400;
    if (rem < 0) { rem += 86400; days -= 1; }
    // civil-from-days (H. Hinnant), proleptic Gregorian
    int64_t z = days + 719468;
    const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
    const unsigned doe = (unsigned)(z - era * 146097);
    const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    int64_t y = (int64_t)yoe + era * 400;
    const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const unsigned mp = (5 * doy + 2) / 153;
    const unsigned d = doy - (153 * mp + 2) / 5 + 1;
    const unsigned m = mp < 10 ? mp + 3 : mp - 9;
    if (m <= 2) y += 1;
    char buf[96];
    int n = snprintf(buf, sizeof buf, "%04lld-%02u-%02u %02lld:%02lld:%02lld", (long long)y, m, d, (long long)(rem / 3600),
                     (long long)(rem % 3600 / 60), (long long)(rem % 60));
    if (nanos != 0) {
        if (nanos % 1000000 == 0) n += snprintf(buf + n, sizeof buf - n, ".%03d", nanos / 1000000);
        else if (nanos % 1000 == 0) n += snprintf(buf + n, sizeof buf - n, ".%06d", nanos / 1000);
        else n += snprintf(buf + n, sizeof buf - n, ".%09d", nanos);
    }
    snprintf(buf + n, sizeof buf - n, " UTC");
    return buf;
}

inline std::string render_table(const std::vector<std::vector<std::string>> &rows) {
    std::vector<size_t> w;
    for (const auto &r : rows) {
        if (w.size() < r.size()) w.resize(r.size(), 0);
        for (size_t c = 0; c < r.size(); c++) w[c] = std::max(w[c], r[c].size());
    }
    std::string sep = "+";
    for (size_t c = 0; c < w.size(); c++) sep += std::string(w[c] + 2, '-') + "+";
    sep += "\n";
    std::string out = sep;
    for (const auto &r : rows) {
        out += "|";
        for (size_t c = 0; c < w.size(); c++) {
            const std::string &cell = c < r.size() ? r[c] : std::string();
            out += " " + cell + std::string(w[c] - cell.size(), ' ') + " |";
        }
        out += "\n" + sep;
    }
    return out;
}

inline std::string render(const Summary &s, const std::vector<PartitionRow> &parts) {
    auto u = [](uint64_t v) { return std::to_string(v); };
    const std::string eq(120, '='), dash(120, '-');
    std::string o = "\n" + eq + "\nCalculating statistics...\n";
    o += "Topic " + s.topic + "\n";
    o += "Scanning took: " + u(s.duration_secs) + " seconds\n";
    o += "Estimated Msg/s: " + u(s.overall_count / (s.duration_secs > 1 ? s.duration_secs : 1)) + "\n";  // main.rs:130
    o += dash + "\nEarliest Message: " + format_utc(s.earliest_s, s.earliest_ns) + "\n";
    o += "Latest Message: " + format_utc(s.latest_s, 0) + "\n" + dash + "\n";
    o += "Largest Message: " + u(s.largest_message) + " bytes\n";
    o += "Smallest Message: " + u(s.smallest_message) + " bytes\n";
    o += "Topic Size: " + u(s.overall_size) + " bytes\n";
    if (s.has_alive_keys) o += dash + "\nAlive keys: " + u(s.alive_keys) + "\n" + dash + "\n";  // main.rs:139-146
    o += eq + "\n";
    std::vector<std::vector<std::string>> rows;
    rows.push_back({"P", "< OS", "> OS", "Total", "Alive", "Tmb", "DR", "K Null", "K !Null", "P-Bytes", "K-Bytes", "V-Bytes",
                    "A K-Sz", "A V-Sz", "A M-Sz"});  // main.rs:150
    for (const auto &p : parts) {
        char dr[64];
        snprintf(dr, sizeof dr, "%.4f", (double)p.dirty_ratio);  // {0:.4}
        rows.push_back({std::to_string(p.partition), std::to_string(p.start_offset), std::to_string(p.end_offset), u(p.total),
                        u(p.alive), u(p.tombstones), dr, u(p.key_null), u(p.key_non_null),
                        u(p.key_size_sum + p.value_size_sum), u(p.key_size_sum), u(p.value_size_sum), u(p.key_size_avg),
                        u(p.value_size_avg), u(p.message_size_avg)});
    }
    o += "| K = Key, V = Value, P = Partition, Tmb = Tombstone(s), Sz = Size\n";
    o += "| DR = Dirty Ratio, A = Average, Lst = last, < OS = start offset, > OS = end offset\n";
    o += render_table(rows);
    o += "\n" + eq + "\n";
    return o;
}

}  // namespace kta_report
