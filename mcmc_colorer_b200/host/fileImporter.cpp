#include "fileImporter.h"

#include <cstdlib>
#include <iostream>
#include <unordered_set>

fileImporter::fileImporter(std::string graphFileName, std::string) {
	std::unordered_set<std::string> names;
	graphFile.open(graphFileName.c_str());
	if (graphFile) std::getline(graphFile, inStr);                 // header line (fileImporter.cpp:24-25)
	else { std::cout << "Error opening graph file" << std::endl; exit(-1); }
	float ww;
	while (graphFile) {                                            // fileImporter.cpp:35-49
		std::getline(graphFile, inStr);
		if (inStr == "") continue;
		ss << inStr; ss >> src; ss >> dst; ss >> ww;
		names.insert(src); names.insert(dst);
		ss.str(""); ss.clear();
	}
	nNodes = (uint32_t)names.size();
	int i = 0;
	for (auto it = names.begin(); it != names.end(); ++it) {        // fileImporter.cpp:54-58
		geneMap.insert(std::pair<std::string, int>(*it, i));
		inverseGeneMap.insert(std::pair<int, std::string>(i, *it));
		i++;
	}
}

fileImporter::~fileImporter() { graphFile.close(); }

void fileImporter::fRewind() {                                     // fileImporter.cpp:73-84
	std::string line;
	graphFile.clear(); graphFile.seekg(0);
	std::getline(graphFile, line);
	ss.str(""); ss.clear();
}

bool fileImporter::getNextEdge() {                                 // fileImporter.cpp:118-143
	edgeIsValid = false;
	if (!graphFile) return false;
	do { std::getline(graphFile, inStr); } while ((inStr == "") && (graphFile));
	if (!graphFile) return false;
	ss << inStr; ss >> src; ss >> dst; ss >> ww_d;
	srcIdx = geneMap.at(src); dstIdx = geneMap.at(dst); edgeWgh = ww_d;
	edgeIsValid = true;
	ss.str(""); ss.clear();
	return true;
}
