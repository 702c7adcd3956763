// report_golden — prints the report for counters given on stdin (no GPU): used by tests/test_report.py to
// pin the formatter (src/main.rs:123-179) to demo_output.png.
// stdin: topic secs count earliest_s earliest_ns latest_s largest smallest size has_alive alive nrows, then per row:
//        P start end total alive tomb key_null key_non_null ksum vsum
#include <iostream>
#include "kta_report.hpp"
int main() {
    kta_report::Summary s{};
    int has = 0, n = 0;
    std::cin >> s.topic >> s.duration_secs >> s.overall_count >> s.earliest_s >> s.earliest_ns >> s.latest_s >> s.largest_message >>
        s.smallest_message >> s.overall_size >> has >> s.alive_keys >> n;
    s.has_alive_keys = has != 0;
    std::vector<kta_report::PartitionRow> rows;
    for (int i = 0; i < n; i++) {
        kta_report::PartitionRow r{};
        std::cin >> r.partition >> r.start_offset >> r.end_offset >> r.total >> r.alive >> r.tombstones >> r.key_null >> r.key_non_null >>
            r.key_size_sum >> r.value_size_sum;
        // same arithmetic as kta_avg / kta_dirty_ratio (src/metric.rs:132-167)
        r.key_size_avg = r.key_size_sum ? r.key_size_sum / r.alive : 0;
        r.value_size_avg = r.value_size_sum ? r.value_size_sum / r.alive : 0;
        r.message_size_avg = (r.key_size_sum + r.value_size_sum) ? (r.key_size_sum + r.value_size_sum) / r.alive : 0;
        r.dirty_ratio = (r.total > 0 && r.tombstones > 0) ? (float)r.tombstones / ((float)r.total / 100.0f) : 0.0f;
        rows.push_back(r);
    }
    std::cout << kta_report::render(s, rows);
}
